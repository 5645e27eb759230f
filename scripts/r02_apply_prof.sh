#!/bin/bash
mkdir -p gpurun_out
STEPS=30 timeout 300 python scripts/retained_bench.py 2>&1 | grep "acceptor 0"
STEPS=12 timeout 900 ncu --set full --clock-control none --import-source on -k regex:apply_list_kernel -s 14 -c 1 -o gpurun_out/apply_full -f python scripts/retained_bench.py > gpurun_out/apply_ncu_full.log 2>&1
tail -2 gpurun_out/apply_ncu_full.log
