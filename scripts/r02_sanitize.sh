# compute-sanitizer over the kernels added in round 2 (union walkers / scheduler / score, SA replay, join expressions,
# k-opt, parallel build_fast_records): memcheck on the new test files, racecheck on a subset
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_union.py tests/test_gpu_k_opt.py tests/test_gpu_join_expr.py -x -q -k "not trajectory and not whole_stream and not default_list_union" 2>&1 | tail -6 > gpurun_out/r02_memcheck.log
echo "memcheck rc=$?" >> gpurun_out/r02_memcheck.log
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_union.py -x -q -k "apply_chain or windows_grow or simulated" 2>&1 | tail -6 > gpurun_out/r02_racecheck.log
echo "racecheck rc=$?" >> gpurun_out/r02_racecheck.log
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_nearby_step.py tests/test_gpu_change_step.py -x -q -k "loop or apply or simulated" 2>&1 | tail -6 >> gpurun_out/r02_racecheck.log
echo "racecheck2 rc=$?" >> gpurun_out/r02_racecheck.log
cat gpurun_out/r02_memcheck.log gpurun_out/r02_racecheck.log
