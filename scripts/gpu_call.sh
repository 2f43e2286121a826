mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_parity_r2.py -x -q -m gpu > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2d_pytest.log
CEV_LS_FORK=0 CEV_LS_SKIP=1 timeout 120 python scripts/time_ls.py
CEV_LS_SKIP=1 CEV_LS_GRID_MEM=74 timeout 120 python scripts/time_ls.py
CEV_LS_SKIP=1 CEV_LS_GRID_MEM=96 timeout 120 python scripts/time_ls.py
timeout 200 python scripts/time_roles.py 2>/dev/null | head -1
for g in "52 96" "64 84" "74 74"; do set -- $g; echo "opp=$1 mem=$2"; CEV_LS_GRID_OPP=$1 CEV_LS_GRID_MEM=$2 timeout 200 python scripts/time_roles.py 2>/dev/null | head -1; done
grep -A4 "teacher" gpurun_out/parity_report.json | grep "max_abs_err"
