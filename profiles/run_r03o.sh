cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ATTWARP_FUZZ_CASES=400 timeout 2400 python -m pytest tests/test_gpu_remap_fuzz.py -m gpu -q -x -n 4 > gpurun_out/r03o_fuzz.log 2>&1; echo "pytest exit $?" >> gpurun_out/r03o_fuzz.log; tail -15 gpurun_out/r03o_fuzz.log
