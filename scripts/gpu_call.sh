mkdir -p gpurun_out
timeout 300 python scripts/quick_k2.py
timeout 900 python -m pytest tests/test_gpu_deepqn.py tests/test_gpu_parity_r2.py -x -q -m gpu 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:deepqn -c 15 --csv --log-file gpurun_out/r2c_k2_launches.csv python scripts/quick_k2.py > /dev/null 2>&1
grep -E "gpu__time_duration" gpurun_out/r2c_k2_launches.csv | head -5 | cut -d, -f5,15
