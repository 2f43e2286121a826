timeout 600 python scripts/stress_roles.py 2>&1 | tail -8
