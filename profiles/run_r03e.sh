cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mask_path.py -m gpu -q -x > gpurun_out/r03e_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r03e_pytest.log; tail -3 gpurun_out/r03e_pytest.log
timeout 600 python profiles/row_kernels_probe.py 2>&1 | grep -v Warning | tee gpurun_out/r03e_row_kernels.txt
