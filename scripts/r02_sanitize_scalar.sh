# compute-sanitizer over the scalar step / loop kernels touched late in round 2 (change_finish_kernel as a template over
# the scoring program with the staged block, tabu_accept_kernel / tabu_record_kernel)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_change_step.py -x -q -k "tabu or stateful or step_matches or loop" 2>&1 | tail -6 > gpurun_out/r02_scalar_memcheck.log
echo "memcheck rc=${PIPESTATUS[0]}" >> gpurun_out/r02_scalar_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_change_step.py -x -q -k "tabu and 5-0-0-0 or stateful" 2>&1 | tail -6 > gpurun_out/r02_scalar_racecheck.log
echo "racecheck rc=${PIPESTATUS[0]}" >> gpurun_out/r02_scalar_racecheck.log
cat gpurun_out/r02_scalar_memcheck.log gpurun_out/r02_scalar_racecheck.log
