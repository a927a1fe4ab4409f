ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/cnn_launches.csv python tests/dev/cnn_once.py 256 > /dev/null 2>&1
grep "conv_in2d" gpurun_out/cnn_launches.csv | cut -d, -f5,15- | head -3
timeout 300 python tests/dev/fixed_cost.py 2>&1 | grep -E "^256|^128 |^16 " | cut -c1-60
