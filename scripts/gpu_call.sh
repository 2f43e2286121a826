mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2b_pytest_all.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2b_pytest_all.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2b_bench_n1.json').read().strip().splitlines()[-1])
print(d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d.get('parity'), d['roofline'].get('frac'), d['gpu_launches'])"
