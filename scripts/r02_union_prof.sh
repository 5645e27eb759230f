#!/bin/bash
mkdir -p gpurun_out
python scripts/union_bench.py 1024 256; SFGPU_NO_COND=1 python scripts/union_bench.py 1024 256
SFGPU_NO_COND=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/union_launches.csv python scripts/union_bench.py 1024 32 > /dev/null 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open('gpurun_out/union_launches.csv', errors='ignore')))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
agg = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) > 14:
        agg[r[4][:80]].append(float(r[14]))
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{len(v):5d} x {sum(v)/len(v)/1e3:7.1f} us = {sum(v)/tot*100:5.1f}%  {k}")
PY
