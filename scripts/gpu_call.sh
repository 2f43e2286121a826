mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/r2_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread --clock-control none --csv --log-file gpurun_out/r2_pop_kernels_ncu.csv python scripts/ncu_pop_kernels.py 1024 2 > /dev/null 2>&1; echo "ncu pop rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:es_perturb_kernel -s 1 -c 1 -o gpurun_out/r2_k5 -f python scripts/ncu_pop_kernels.py 1024 2 > /dev/null 2>&1
ncu -i gpurun_out/r2_k5.ncu-rep --page raw --csv > gpurun_out/r2_k5_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:deepqn_fc_tc -s 1 -c 1 -o gpurun_out/r2_k2_fc3 -f python scripts/quick_k2.py > /dev/null 2>&1
ncu -i gpurun_out/r2_k2_fc3.ncu-rep --page raw --csv > gpurun_out/r2_k2_fc3_raw.csv 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:deepqn -c 15 --csv --log-file gpurun_out/r2_k2_launches.csv python scripts/quick_k2.py > /dev/null 2>&1
ls gpurun_out | grep r2_ | tr '\n' ' '
