cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ATTWARP_FUZZ_CASES=400 timeout 2400 python -m pytest tests/test_gpu_remap_fuzz.py -m gpu -q -n 4 > gpurun_out/r03q_fuzz.log 2>&1; echo "pytest exit $?" >> gpurun_out/r03q_fuzz.log; tail -12 gpurun_out/r03q_fuzz.log | cut -c1-300
