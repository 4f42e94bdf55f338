cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ATTWARP_FUZZ_CASES=400 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_remap_fuzz.py -m gpu -q -x -k "uniform_batch_random and 62" > gpurun_out/r03p_memcheck62.log 2>&1; echo "exit $?" >> gpurun_out/r03p_memcheck62.log
grep -n "Invalid\|at aw::\|at void aw\|by thread\|Address\|in \/" gpurun_out/r03p_memcheck62.log | head -30
tail -5 gpurun_out/r03p_memcheck62.log
