timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r2f.json')); print(d['value'], d['ms_per_step'], d['config']['stage_ms'], d['config4']['ms_per_step'], d['config4']['split_ms_eager'], d['config3']['ms_per_step'], d['config5']['frames_128']['ms_per_step'])"
tail -3 gpurun_out/bench_r2f.err
