timeout 900 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_parity_r2.py -x -q -m gpu 2>&1 | tail -2
CEV_LS_SKIP=2 CEV_LS_GRID_OPP=52 timeout 120 python scripts/time_ls.py
for i in 1 2; do timeout 200 python scripts/time_roles.py 2>/dev/null | head -1 | cut -d: -f2 | cut -d, -f1; done
