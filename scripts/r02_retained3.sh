#!/bin/bash
mkdir -p gpurun_out
STEPS=12 timeout 900 ncu --set full --clock-control none --import-source on -k regex:nearby_step_cached -s 14 -c 1 -o gpurun_out/retained_full -f python scripts/retained_bench.py > gpurun_out/retained_ncu_full.log 2>&1
tail -3 gpurun_out/retained_ncu_full.log
ls -la gpurun_out/*.ncu-rep
