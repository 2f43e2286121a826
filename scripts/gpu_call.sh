timeout 600 python -m pytest tests/test_gpu_rollout.py -x -q -m gpu -k "extreme" 2>&1 | tail -15
