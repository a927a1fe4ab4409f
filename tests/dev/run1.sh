timeout 120 python tests/dev/fixed_cost.py 2>&1 | cut -c1-80
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
