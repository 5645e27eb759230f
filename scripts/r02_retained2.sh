#!/bin/bash
mkdir -p gpurun_out
STEPS=20 timeout 300 python scripts/retained_bench.py > gpurun_out/retained_bench.log 2>&1
tail -3 gpurun_out/retained_bench.log
STEPS=6 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/retained_launches.csv python scripts/retained_bench.py > gpurun_out/retained_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/retained_launches.csv', errors='ignore')))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
agg = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) > 14:
        agg[r[4][:70]].append(float(r[14]))
for k, v in agg.items():
    print(f"{len(v):4d} x {sum(v)/len(v)/1e3:8.1f} us  {k}")
PY
