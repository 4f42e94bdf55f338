cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ATTWARP_FUZZ_CASES=400 timeout 3000 python -m pytest tests/test_gpu_stage_fuzz.py -m gpu -q -n 6 -k torch_helpers > gpurun_out/r03v_fuzz.log 2>&1; echo "pytest exit $?" >> gpurun_out/r03v_fuzz.log; tail -14 gpurun_out/r03v_fuzz.log | cut -c1-330
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r03v_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r03v_pytest.log; tail -4 gpurun_out/r03v_pytest.log
