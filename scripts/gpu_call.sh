for n in 2 3 2 3; do echo -n "nstreams=$n: "; CEV_LS_NSTREAMS=$n timeout 200 python scripts/time_roles.py 2>/dev/null | head -1 | cut -d: -f2 | cut -d, -f1; done
timeout 600 python -m pytest tests/test_gpu_rollout.py -x -q -m gpu -k "roles or lockstep" 2>&1 | tail -2
