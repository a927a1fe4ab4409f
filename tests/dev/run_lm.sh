for v in 0 1; do
SEDB_LIB_PATH=$PWD/tests/dev/lib_v$v.so timeout 300 python tests/dev/lm_time.py 256 2>&1 | tail -1
done
SEDB_LIB_PATH=$PWD/tests/dev/lib_v1.so timeout 300 python tests/dev/phase_prof.py 256
