mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_parity_r2.py -x -q -m gpu > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2d_pytest.log
CEV_LS_FORK=0 CEV_LS_SKIP=2 timeout 120 python scripts/time_ls.py
CEV_LS_SKIP=2 CEV_LS_GRID_OPP=52 timeout 120 python scripts/time_ls.py
timeout 200 python scripts/time_roles.py 2>/dev/null | head -1
for g in "43 105" "37 111" "52 96"; do set -- $g; echo "opp=$1 mem=$2"; CEV_LS_GRID_OPP=$1 CEV_LS_GRID_MEM=$2 timeout 200 python scripts/time_roles.py 2>/dev/null | head -1; done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value']/1e6, d['ms_per_step'])"
