mkdir -p gpurun_out
for l2p in 256 128 64; do for skip in 0 1; do CEV_LS_L2P=$l2p CEV_LS_SKIP=$skip timeout 120 python scripts/time_ls.py 1024x16 8192x1 2>&1 | sed "s/^/l2p=$l2p /"; done; done
