ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_train.csv python tests/dev/train_once.py > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_train.csv')) if len(r)>10 and r[0].isdigit()]
# last step = after the last adam_prepare before final
names=[r[4] for r in rows]
idx=[i for i,n in enumerate(names) if 'adam_amsgrad_dev' in n]
seg=rows[idx[-2]+1:idx[-1]+1]
tot=0
for r in seg:
    n=r[4].split('(')[0].replace('void ','').replace('sedb::','')
    print("  %-34s grid %-14s %8.2f us" % (n[:34], r[8], float(r[-1])/1000)); tot+=float(r[-1])/1000
print("  launches", len(seg), "sum", tot)
PY
