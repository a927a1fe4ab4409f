set -x
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4
timeout 300 python tests/dev/lm_time.py 256
timeout 900 python bench.py > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r2d.json')); print(d['value'], d['ms_per_step'], d['config']['stage_ms'], d['roofline']['frac'], d['e2e']['value'], d['config4']['ms_per_step'], d['config3']['ms_per_step'])"
ncu --set full --clock-control none --import-source on -k regex:logmel_fused -s 3 -c 1 -o gpurun_out/logmel_r2 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-pcm16 --no-configs --e2e-clips 8 > /dev/null 2> gpurun_out/ncu2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-clips 8 > gpurun_out/bench_under_ncu.json 2> gpurun_out/ncu1.err
