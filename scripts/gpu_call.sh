mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_parity_r2.py -m gpu -q -x 2>&1 | tail -15
for mode in tc fp32; do for skip in 0 1; do CEV_LS_MEMBER=$mode CEV_LS_SKIP=$skip timeout 120 python scripts/time_ls.py 1024x16 4096x16 2>&1 | sed "s/^/member=$mode /"; done; done
CEV_MT_L2P=128 CEV_LS_MEMBER=tc CEV_LS_SKIP=1 timeout 120 python scripts/time_ls.py 1024x16 2>&1 | sed "s/^/l2p128 member=tc /"
timeout 120 python scripts/ab_step.py 2>&1 | tail -4
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed --clock-control none -k regex:"ls_l1|ls_member_tc" -s 10 -c 4 --csv --log-file gpurun_out/r2_mtc_ncu.csv python scripts/prof_ls.py 1024 16 1 > /dev/null 2>&1; grep -E "ls_" gpurun_out/r2_mtc_ncu.csv | cut -d, -f5,13- | head -16
