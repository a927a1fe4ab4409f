timeout 600 python -m pytest tests/test_gpu_cnn.py tests/test_gpu_train_native.py tests/test_gpu_robustness.py -x -q 2>&1 | tail -4
timeout 120 python tests/dev/fixed_cost.py 2>&1 | cut -c1-90
bash tests/dev/run2.sh 2>/dev/null | head -0
for B in 256; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_b$B.csv python tests/dev/cnn_once.py $B > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_b$B.csv')) if len(r)>10 and r[0].isdigit()]
print("B=$B")
for r in rows[-9:]:
    print("  ", r[4][:34], r[8], float(r[-1])/1000)
print("  sum", sum(float(r[-1]) for r in rows[-9:])/1000)
PY
done
