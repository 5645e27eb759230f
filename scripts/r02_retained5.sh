#!/bin/bash
for c in 1 2 3 4 6 8; do
  echo "== chunks $c"
  SFGPU_NB_CHUNKS=$c STEPS=30 timeout 300 python scripts/retained_bench.py 2>&1 | grep "acceptor 0"
done
