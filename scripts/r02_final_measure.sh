#!/bin/bash
# round-2 measurement pass: the driver's default invocation, its ncu launch list, one full capture of the headline kernel
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
tail -c 300 gpurun_out/r02_bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 6 --warmup 3 --loop-steps 0 --no-extra > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:score_list_change_fast_kernel -s 6 -c 1 -f -o gpurun_out/prof_r02_fast python bench.py --steps 6 --warmup 3 --loop-steps 0 --no-extra > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:nearby_step_kernel -s 2 -c 1 -f -o gpurun_out/prof_r02_nearby python bench.py --steps 6 --warmup 3 --loop-steps 0 --no-extra > /dev/null 2>&1
ls -la gpurun_out | tail -8
