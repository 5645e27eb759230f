#!/bin/bash
# round-end measurement pass: bench lines, ncu launch list, one full capture of the headline kernel
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_final_cvrp.json 2> gpurun_out/bench_final_cvrp.err
python bench.py --workload graph_coloring --replicas 192 --steps 50 --warmup 3 --loop-steps 0 > gpurun_out/bench_final_gc.json 2>/dev/null
python bench.py --workload job_shop --replicas 200 --steps 50 --warmup 3 --loop-steps 0 > gpurun_out/bench_final_js.json 2>/dev/null
python bench.py --replicas 1 --distinct 1 --steps 200 --warmup 5 > gpurun_out/bench_final_R1.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 6 --warmup 3 --loop-steps 0 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:score_list_change_fast_kernel -s 6 -c 1 -f -o gpurun_out/prof_fast_final python bench.py --steps 6 --warmup 3 --loop-steps 0 > /dev/null 2>&1
ls -la gpurun_out
