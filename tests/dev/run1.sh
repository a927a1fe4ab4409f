timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_train_native.py -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 5 --no-cpu-baseline --no-pcm16 --e2e-clips 16 2> gpurun_out/bench_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms',d['ms_per_step'],d['config']['stage_ms'])
for k in ('config3','config4','config5'):
    print(k, json.dumps({a:b for a,b in d.get(k).items() if a not in ('workload','roofline')})[:900])
"
tail -5 gpurun_out/bench_err.log
