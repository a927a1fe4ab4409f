timeout 900 python -m pytest tests/test_gpu_logmel.py tests/test_gpu_pcm16.py tests/test_gpu_callers.py tests/test_gpu_robustness.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python tests/dev/lm_time.py 256
timeout 300 python tests/dev/lm_time.py 256
python tests/dev/inv_check.py | awk '{print $2,$3,$4,$6}' | tr '\n' ';'
