mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2g_pytest_all.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2g_pytest_all.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6)"
