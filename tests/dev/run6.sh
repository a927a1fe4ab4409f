set -x
ncu --set full --clock-control none --import-source on -k regex:logmel_fused -s 3 -c 1 -o gpurun_out/logmel_r2 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-pcm16 --no-configs --e2e-clips 8 > /dev/null 2> gpurun_out/ncu2.err
ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 14 -c 7 -o gpurun_out/conv_r2 -f python tests/dev/cnn_once.py 256 > /dev/null 2> gpurun_out/ncu3.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-clips 8 > gpurun_out/bench_under_ncu.json 2> gpurun_out/ncu1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_train.csv python tests/dev/train_once.py > /dev/null 2> gpurun_out/ncu4.err
timeout 900 compute-sanitizer --tool memcheck python tests/dev/sanitize_small.py > gpurun_out/sanitizer_memcheck_r2.txt 2>&1
timeout 900 compute-sanitizer --tool synccheck python tests/dev/sanitize_small.py > gpurun_out/sanitizer_synccheck_r2.txt 2>&1
tail -n 2 gpurun_out/sanitizer_memcheck_r2.txt gpurun_out/sanitizer_synccheck_r2.txt
ls -la gpurun_out/*.ncu-rep | tail -3
