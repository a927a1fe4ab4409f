timeout 900 python -m pytest tests/test_gpu_resample.py -q -m gpu 2>&1 | tail -2
timeout 300 python tests/dev/resample_time.py 2>&1 | tail -4
