timeout 600 python scripts/stress_roles.py 2>&1 | tail -6
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for i in 1 2; do timeout 200 python scripts/time_roles.py 2>/dev/null | head -1 | cut -d: -f2 | cut -d, -f1; done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value']/1e6, d['ms_per_step'])"
