python -m pytest tests/test_gpu_cnn.py -q -m gpu 2>&1 | tail -3
