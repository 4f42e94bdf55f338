cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ATTWARP_FUZZ_CASES=400 timeout 3000 python -m pytest tests/test_gpu_pipeline_fuzz.py -m gpu -q -n 6 > gpurun_out/r03w_fuzz.log 2>&1; echo "pytest exit $?" >> gpurun_out/r03w_fuzz.log; tail -14 gpurun_out/r03w_fuzz.log | cut -c1-330
