#!/bin/bash
# round-2 final evidence: GPU tests, smoke, bench line, launch list of the bench command, full captures of the kernels
# behind the committed e2e step
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r02_gpu_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r02_gpu_tests.log
tail -3 gpurun_out/r02_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
tail -2 gpurun_out/r02_bench_final.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --loop-steps 0 > gpurun_out/r02_bench_under_ncu.log 2>&1
STEPS=12 timeout 900 ncu --set full --clock-control none --import-source on -k regex:nearby_regen -s 14 -c 1 -o gpurun_out/regen_full -f python scripts/retained_bench.py > /dev/null 2>&1
STEPS=12 timeout 900 ncu --set full --clock-control none --import-source on -k regex:nearby_step_cached -s 14 -c 1 -o gpurun_out/retained_full -f python scripts/retained_bench.py > /dev/null 2>&1
STEPS=12 timeout 900 ncu --set full --clock-control none --import-source on -k regex:nearby_finish -s 30 -c 1 -o gpurun_out/finish_full -f python scripts/retained_bench.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_final.json"))
print("value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
print("device_loop", d["device_loop"]["ms_per_step"], d["device_loop"]["moves_evaluated_per_s"])
print("default_search", d["default_search"]["ms_per_step"], d["default_search"]["moves_evaluated_per_s"], d["default_search"]["scored_over_evaluated"])
for k, v in d["extra"].items():
    print(k, v["value"], v["roofline"]["frac"], v["e2e"]["value"], v["e2e"]["ms_per_step"], v.get("default_search", {}).get("ms_per_step"))
PY
