#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 2 4 6; do
  echo "== SFGPU_NBC_DEBUG=$dbg"
  SFGPU_NBC_DEBUG=$dbg STEPS=30 timeout 300 python scripts/retained_bench.py 2>&1 | grep "acceptor 0"
done
