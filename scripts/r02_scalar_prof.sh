#!/bin/bash
mkdir -p gpurun_out
for w in job_shop graph_coloring; do
  python scripts/scalar_step_bench.py $w
  STEPS=5 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/scalar_launches_$w.csv python scripts/scalar_step_bench.py $w > /dev/null 2>&1
  python - <<PY
import csv, collections
rows = list(csv.reader(open('gpurun_out/scalar_launches_$w.csv', errors='ignore')))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
agg = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) > 14:
        agg[(r[4][:100], r[7], r[8])].append(float(r[14]))
for k, v in agg.items():
    print(f"{len(v):4d} x {sum(v)/len(v)/1e3:8.1f} us  {k}")
PY
done
