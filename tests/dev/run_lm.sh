timeout 1500 python -m pytest tests/test_gpu_train.py tests/test_gpu_train_native.py -q -m gpu -x 2>&1 | tail -3
timeout 900 python bench.py --no-cpu-baseline --no-pcm16 --e2e-clips 8 > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r2h.json')); print(d['value'], d['config4']['ms_per_step'], d['config4']['split_ms_eager'])"; tail -2 gpurun_out/bench_r2h.err
