timeout 600 python tests/dev/fuzz_logmel.py 1 2>&1 | tail -5
timeout 600 python tests/dev/fuzz_logmel.py 2 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_logmel.py tests/test_gpu_pcm16.py -q -m gpu 2>&1 | tail -2
timeout 300 python tests/dev/lm_time.py 256 2>&1 | tail -1
