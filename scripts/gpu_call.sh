mkdir -p gpurun_out
for PF in 0 2 4 8; do for skip in 0 1; do COEVONET_LIB=$PWD/coevonet_b200/csrc/var/libcev_pf$PF.so CEV_LS_SPLIT=0 CEV_LS_SKIP=$skip timeout 120 python scripts/time_ls.py 1024x16 8192x1 2>&1 | sed "s/^/pf=$PF /"; done; done
