timeout 300 python tests/dev/lm_time.py 256 2>&1 | tail -1
sleep 2
SEDB_LIB_PATH=$PWD/tests/dev/lib_z23.so timeout 300 python tests/dev/lm_time.py 256 2>&1 | tail -1
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3
