timeout 600 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_parity_r2.py -x -q -m gpu 2>&1 | tail -2
for p in 0 1 2; do echo "prio=$p"; CEV_LS_PRIO=$p timeout 200 python scripts/time_roles.py 2>/dev/null | head -1; done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value']/1e6, d['ms_per_step'])"
