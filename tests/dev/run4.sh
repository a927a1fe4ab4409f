set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-clips 8 > gpurun_out/bench_under_ncu.json 2> gpurun_out/ncu1.err
ncu --set full --clock-control none --import-source on -k regex:logmel_fused -s 3 -c 1 -o gpurun_out/logmel_r2 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-pcm16 --no-configs --e2e-clips 8 > /dev/null 2> gpurun_out/ncu2.err
timeout 900 compute-sanitizer --tool memcheck python tests/dev/sanitize_small.py > gpurun_out/sanitizer_memcheck_r2.txt 2>&1
timeout 900 compute-sanitizer --tool synccheck python tests/dev/sanitize_small.py > gpurun_out/sanitizer_synccheck_r2.txt 2>&1
tail -3 gpurun_out/sanitizer_memcheck_r2.txt gpurun_out/sanitizer_synccheck_r2.txt
ls -la gpurun_out/*.ncu-rep | tail -3
