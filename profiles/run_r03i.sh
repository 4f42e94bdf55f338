cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r03i_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r03i_pytest.log
tail -5 gpurun_out/r03i_pytest.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_remap_edges.py -m gpu -q -x -k "minification or alignment" > gpurun_out/r03i_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r03i_memcheck.log
tail -3 gpurun_out/r03i_memcheck.log
