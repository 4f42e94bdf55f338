# round 2: LANE mapping (conflict-free loads + in-warp transpose) vs QUAD mapping
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_remap_edges.py tests/test_gpu_ragged.py tests/test_gpu_fused_batch.py tests/test_gpu_numpy_path.py -m gpu -q -x > gpurun_out/r02f_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02f_pytest.log
tail -12 gpurun_out/r02f_pytest.log
ATTWARP_QUAD_MAP=2 timeout 900 python -m pytest tests/test_gpu_remap_edges.py tests/test_gpu_ragged.py tests/test_gpu_fused_batch.py tests/test_gpu_numpy_path.py -m gpu -q -x > gpurun_out/r02f_pytest_lane.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02f_pytest_lane.log
tail -5 gpurun_out/r02f_pytest_lane.log
run() { echo "== $1" >> gpurun_out/r02f_probe.txt; env $1 timeout 300 python profiles/s5_probe.py --only "$2" >> gpurun_out/r02f_probe.txt 2>&1; }
run "ATTWARP_QUAD_MAP=0" ""
run "ATTWARP_QUAD_MAP=1" c
run "ATTWARP_QUAD_MAP=2" c
cat gpurun_out/r02f_probe.txt
timeout 300 python profiles/c4_probe.py > gpurun_out/r02f_c4.txt 2>&1; cat gpurun_out/r02f_c4.txt
