timeout 900 compute-sanitizer --tool memcheck python tests/dev/sanitize_small.py > gpurun_out/sanitizer_memcheck_r2.txt 2>&1
timeout 900 compute-sanitizer --tool synccheck python tests/dev/sanitize_small.py > gpurun_out/sanitizer_synccheck_r2.txt 2>&1
tail -n 2 gpurun_out/sanitizer_memcheck_r2.txt gpurun_out/sanitizer_synccheck_r2.txt
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -2
