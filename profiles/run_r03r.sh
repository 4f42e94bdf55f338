cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ATTWARP_FUZZ_CASES=1600 timeout 2400 python -m pytest tests/test_gpu_remap_fuzz.py -m gpu -q -n 6 > gpurun_out/r03r_fuzz.log 2>&1; echo "pytest exit $?" >> gpurun_out/r03r_fuzz.log; tail -12 gpurun_out/r03r_fuzz.log | cut -c1-300
for pol in 1 2; do ATTWARP_QUAD_MAP=$pol ATTWARP_FUZZ_CASES=300 timeout 1200 python -m pytest tests/test_gpu_remap_fuzz.py -m gpu -q -n 6 -k "uniform or ragged" > gpurun_out/r03r_fuzz_map$pol.log 2>&1; echo "pytest exit $?" >> gpurun_out/r03r_fuzz_map$pol.log; tail -3 gpurun_out/r03r_fuzz_map$pol.log | cut -c1-300; done
ATTWARP_FUZZ_CASES=160 timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_remap_fuzz.py -m gpu -q -x > gpurun_out/r03r_fuzz_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r03r_fuzz_memcheck.log; tail -4 gpurun_out/r03r_fuzz_memcheck.log | cut -c1-300
