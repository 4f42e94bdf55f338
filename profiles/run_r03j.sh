cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_torch_path.py tests/test_c5_marginalnet.py tests/test_gpu_autograd.py -m gpu -q -x > gpurun_out/r03j_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r03j_pytest.log; tail -3 gpurun_out/r03j_pytest.log
timeout 600 python profiles/row_kernels_probe.py 2>&1 | grep -v Warning | head -3 | tee gpurun_out/r03j_row_kernels.txt
