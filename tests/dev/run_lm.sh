for v in 0 1 0 1; do SEDB_LIB_PATH=$PWD/tests/dev/lib_e$v.so timeout 300 python tests/dev/lm_time.py 256 2>&1 | tail -1; sleep 2; done
SEDB_LIB_PATH=$PWD/tests/dev/lib_e1.so timeout 900 python -m pytest tests/test_gpu_logmel.py tests/test_gpu_pcm16.py -x -q -m gpu 2>&1 | tail -3
SEDB_LIB_PATH=$PWD/tests/dev/lib_e1.so timeout 300 python tests/dev/phase_prof.py 256 2>/dev/null | head -13
