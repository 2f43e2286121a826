mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/r2e_bench_n1.err
python -c "
import json; d=json.loads(open('gpurun_out/r2e_bench_n1.json').read().strip().splitlines()[-1])
print(d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['parity']['member_match_frac'], d['roofline']['frac'], d['roofline']['whole_step']['frac_of_hbm_peak'], d['gpu_launches'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2e_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/r2e_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
CEV_LS_FORK=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:ls_opp_kernel -s 5 -c 1 -o gpurun_out/r2e_opp -f python scripts/prof_ls.py 1024 16 1 > gpurun_out/r2e_ncu3.log 2>&1; echo "ncu opp rc=$?"
ncu -i gpurun_out/r2e_opp.ncu-rep --page raw --csv > gpurun_out/r2e_opp_raw.csv 2>/dev/null
