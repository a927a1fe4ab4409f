set -x
for v in 0 1 2; do SEDB_LIB_PATH=$PWD/tests/dev/lib_w$v.so timeout 300 python tests/dev/lm_time.py 256 2>&1 | tail -1; done
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8
timeout 300 python tests/dev/dbg_hdr.py 2>&1 | tail -3
