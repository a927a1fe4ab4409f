timeout 600 python -m pytest tests/test_gpu_resample.py -x -q -m gpu 2>&1 | tail -12
timeout 300 python tests/dev/resample_time.py 2>&1 | tail -4
