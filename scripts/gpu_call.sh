for i in 1 2 3; do timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -1; done
for i in 1 2; do timeout 600 python scripts/stress_roles.py 2>&1 | tail -1; done
timeout 600 python -m pytest tests/test_gpu_drivers.py tests/test_gpu_parity_r2.py -x -q -m gpu -p no:randomly 2>&1 | tail -1
