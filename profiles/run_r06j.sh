# round 2, pass r06j: tensor-core LANCZOS parity at explicit sizes; memcheck of the new kernels
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
P=gpurun_out/r06j
timeout 900 python -m pytest tests/test_mask_path.py -m gpu -q > ${P}_pytest_mask.log 2>&1; echo "pytest exit $?" >> ${P}_pytest_mask.log; tail -n 12 ${P}_pytest_mask.log | cut -c1-600
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_mask_path.py tests/test_gpu_numpy_path.py -m gpu -q -x -k "lanczos or mota or marginals or transform" > ${P}_memcheck.log 2>&1; echo "memcheck exit $?" >> ${P}_memcheck.log; tail -n 4 ${P}_memcheck.log | cut -c1-300
