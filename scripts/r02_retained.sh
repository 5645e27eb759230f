#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nearby_retained.py tests/test_gpu_nearby_step.py -x -q -m gpu > gpurun_out/retained_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/retained_tests.log
tail -15 gpurun_out/retained_tests.log
bash scripts/r02_retained2.sh
