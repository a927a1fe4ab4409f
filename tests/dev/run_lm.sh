for v in 1 8 1 8; do SEDB_LIB_PATH=$PWD/tests/dev/lib_m$v.so timeout 300 python tests/dev/lm_time.py 256 2>&1 | tail -1; sleep 2; done
SEDB_LIB_PATH=$PWD/tests/dev/lib_m8.so timeout 300 python tests/dev/phase_prof.py 256 2>/dev/null | tail -16
