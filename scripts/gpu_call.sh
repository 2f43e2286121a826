echo "== 4 issuing warps per CTA"; ./scripts/probe/bw_probe 1024 4
echo "== 1 issuing warp per CTA"; ./scripts/probe/bw_probe 1024 1
