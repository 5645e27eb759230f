#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_change_step.py tests/test_gpu_spec.py -x -q -m gpu 2>&1 | tail -3
for w in graph_coloring job_shop; do
  timeout 600 python bench.py --workload $w --loop-steps 0 > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_$w.json"))
print("$w", "value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
PY
done
