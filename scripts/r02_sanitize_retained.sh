# compute-sanitizer over the retained nearby neighbourhood (nearby_step_cached_kernel, nearby_regen_kernel, the cached
# finish path, apply_list_kernel's protocol words and shared-memory element staging): memcheck + racecheck + synccheck
mkdir -p gpurun_out
K="equals_full_regeneration and (180-11-1 or 40-3 or 30-1) or other_writers or device_loop_with"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_nearby_retained.py -x -q -k "$K" 2>&1 | tail -6 > gpurun_out/r02_retained_memcheck.log
echo "memcheck rc=${PIPESTATUS[0]}" >> gpurun_out/r02_retained_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_nearby_retained.py -x -q -k "180-11-40 or 30-1 or other_writers" 2>&1 | tail -6 > gpurun_out/r02_retained_racecheck.log
echo "racecheck rc=${PIPESTATUS[0]}" >> gpurun_out/r02_retained_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_nearby_retained.py -x -q -k "40-3" 2>&1 | tail -6 > gpurun_out/r02_retained_synccheck.log
echo "synccheck rc=${PIPESTATUS[0]}" >> gpurun_out/r02_retained_synccheck.log
cat gpurun_out/r02_retained_memcheck.log gpurun_out/r02_retained_racecheck.log gpurun_out/r02_retained_synccheck.log
