python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r2i.json')); print(d['value'], d['ms_per_step'], d['config']['stage_ms'], d['roofline']['frac'], d['roofline'].get('traffic_stale'), d['e2e']['value'], d['config3']['ms_per_step'], d['config4']['ms_per_step'], d['config5']['frames_128']['ms_per_step'], d['config5']['frames_1024']['ms_per_step'], d['cpu_baseline']['value'], d['clocks'])"
