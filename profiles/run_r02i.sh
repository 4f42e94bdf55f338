# round 2: software store path for odd widths (word stores everywhere), early strip setup; tests + probes + c4
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_remap_edges.py tests/test_gpu_ragged.py tests/test_gpu_fused_batch.py tests/test_gpu_numpy_path.py tests/test_save_warped_image.py tests/test_mask_path.py tests/test_image_io.py -m gpu -q -x > gpurun_out/r02i_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02i_pytest.log
tail -12 gpurun_out/r02i_pytest.log
ATTWARP_QUAD_MAP=2 timeout 900 python -m pytest tests/test_gpu_remap_edges.py tests/test_gpu_ragged.py tests/test_gpu_numpy_path.py -m gpu -q -x > gpurun_out/r02i_pytest_lane.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02i_pytest_lane.log
tail -4 gpurun_out/r02i_pytest_lane.log
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_remap_edges.py tests/test_gpu_ragged.py -m gpu -q -x > gpurun_out/r02i_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r02i_memcheck.log
tail -4 gpurun_out/r02i_memcheck.log
timeout 300 python profiles/s5_probe.py --only c > gpurun_out/r02i_probe.txt 2>&1; cat gpurun_out/r02i_probe.txt
timeout 300 python profiles/c4_probe.py > gpurun_out/r02i_c4.txt 2>&1; cat gpurun_out/r02i_c4.txt
