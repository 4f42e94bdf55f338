# round 2, pass r08f: stage fuzz at a raised case count (the LANCZOS cases reach the tensor-core kernels at random shapes), memcheck on a slice
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
P=gpurun_out/r08f
ATTWARP_FUZZ_CASES=3000 timeout 1500 python -m pytest tests/test_gpu_stage_fuzz.py -m gpu -q -n 6 -k "lanczos" > ${P}_fuzz_lanczos.log 2>&1; echo "pytest exit $?" >> ${P}_fuzz_lanczos.log; tail -n 3 ${P}_fuzz_lanczos.log | cut -c1-400
ATTWARP_FUZZ_CASES=200 timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_stage_fuzz.py -m gpu -q -x -k "lanczos" > ${P}_fuzz_lanczos_memcheck.log 2>&1; echo "memcheck exit $?" >> ${P}_fuzz_lanczos_memcheck.log; tail -n 3 ${P}_fuzz_lanczos_memcheck.log | cut -c1-300
ATTWARP_FUZZ_CASES=100 timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_stage_fuzz.py -m gpu -q -x -k "lanczos" > ${P}_fuzz_lanczos_racecheck.log 2>&1; echo "racecheck exit $?" >> ${P}_fuzz_lanczos_racecheck.log; tail -n 3 ${P}_fuzz_lanczos_racecheck.log | cut -c1-300
