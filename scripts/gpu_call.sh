mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/r2g_bench_n1.err
python -c "
import json; d=json.loads(open('gpurun_out/r2g_bench_n1.json').read().strip().splitlines()[-1])
print(d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['parity']['member_match_frac'], d['roofline']['frac'], d['roofline']['whole_step']['frac_of_hbm_peak'], d['gpu_launches'], d['roofline']['second_kernel']['us_per_launch'], d['roofline']['us_per_launch'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2g_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/r2g_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
