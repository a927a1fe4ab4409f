timeout 900 python -m pytest tests/test_gpu_logmel.py tests/test_gpu_pcm16.py tests/test_gpu_callers.py -x -q -m gpu 2>&1 | tail -3
for v in a b a b; do if [ $v = a ]; then SEDB_LIB_PATH=$PWD/tests/dev/lib_mv0.so timeout 300 python tests/dev/lm_time.py 256 2>&1 | tail -1; else timeout 300 python tests/dev/lm_time.py 256 2>&1 | tail -1; fi; sleep 2; done
timeout 300 python tests/dev/phase_prof.py 256 2>/dev/null | head -13
