ncu --set full --clock-control none --import-source on -k regex:logmel_fused -s 3 -c 1 -o gpurun_out/logmel_r2 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-pcm16 --no-configs --e2e-clips 8 > /dev/null 2> gpurun_out/ncu2.err
ls -la gpurun_out/logmel_r2.ncu-rep
timeout 600 python -m pytest tests/test_gpu_logmel.py -q -m gpu 2>&1 | tail -1
