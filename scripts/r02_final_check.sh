#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r02_gpu_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r02_gpu_tests.log
tail -3 gpurun_out/r02_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
tail -2 gpurun_out/r02_bench_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
tail -c 400 gpurun_out/r02_bench_reference_arm.json
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_final.json"))
print("value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "launches", d["gpu_launches"], d["clocks"])
print("device_loop", d["device_loop"]["ms_per_step"], d["device_loop"]["moves_evaluated_per_s"])
print("default_search", d["default_search"]["ms_per_step"], d["default_search"]["moves_evaluated_per_s"], d["default_search"]["scored_over_evaluated"])
for k, v in d["extra"].items():
    print(k, v["value"], v["roofline"]["frac"], v["e2e"]["value"], v["e2e"]["ms_per_step"], v.get("default_search", {}).get("ms_per_step"))
PY
