python tests/dev/lm_sustain.py 2>&1 | tail -14
