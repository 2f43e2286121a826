mkdir -p gpurun_out
run() { echo "== TC=$1 opp=$2 mem=$3"; CEV_LS_MEMBER_TC=$1 CEV_LS_GRID_OPP=$2 CEV_LS_GRID_MEM=$3 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value']/1e6, d['ms_per_step'], d.get('parity',{}).get('member_match_frac'))"; }
run 0 0 0
run 1 0 0
run 1 64 84
run 1 74 74
run 1 52 96
run 1 84 64
