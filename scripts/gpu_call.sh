run() { echo -n "== $*: "; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value']/1e6, d['ms_per_step'])"; }
run CEV_LS_ALT=1
run CEV_LS_GRID_OPP=43 CEV_LS_GRID_MEM=103
timeout 300 python scripts/ab_step.py 2>&1 | tail -12
timeout 300 python scripts/step_breakdown.py 2>&1 | tail -12
