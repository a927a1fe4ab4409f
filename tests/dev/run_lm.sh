timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3
timeout 300 python tests/dev/fixed_cost.py 2>&1 | grep -E "^256|^128 |^16 |^1 |m5" | cut -c1-60
