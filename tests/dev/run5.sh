set -x
timeout 900 python bench.py > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err
sleep 3
ncu --set full --clock-control none --import-source on -k regex:logmel_fused -s 3 -c 1 -o gpurun_out/logmel_r2 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-pcm16 --no-configs --e2e-clips 8 > /dev/null 2> gpurun_out/ncu2.err
ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 14 -c 7 -o gpurun_out/conv_r2 -f python tests/dev/cnn_once.py 256 > /dev/null 2> gpurun_out/ncu3.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-clips 8 > gpurun_out/bench_under_ncu.json 2> gpurun_out/ncu1.err
timeout 900 compute-sanitizer --tool memcheck python tests/dev/sanitize_small.py > gpurun_out/sanitizer_memcheck_r2.txt 2>&1
timeout 900 compute-sanitizer --tool synccheck python tests/dev/sanitize_small.py > gpurun_out/sanitizer_synccheck_r2.txt 2>&1
SEDB_LIB_PATH=$PWD/tests/dev/lib_bf16.so timeout 900 python -m pytest tests/test_gpu_logmel.py tests/test_gpu_pcm16.py -q -m gpu > gpurun_out/pytest_gpu_bf16.log 2>&1
tail -2 gpurun_out/pytest_gpu_bf16.log gpurun_out/sanitizer_memcheck_r2.txt gpurun_out/sanitizer_synccheck_r2.txt
python -c "
import json; d=json.load(open('gpurun_out/bench_r2e.json')); print(d['value'], d['ms_per_step'], d['config']['stage_ms'], d['roofline']['frac'], d['e2e']['value'], d['config4']['ms_per_step'], d['config3']['ms_per_step'], d['clocks'])"
