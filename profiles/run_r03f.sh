# round 2: unaligned rows with a one-pixel halo per warp (lane-local funnel shifts + one shuffle, byte stores only at row ends)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for pol in 0 1 2; do
ATTWARP_QUAD_MAP=$pol timeout 1200 python -m pytest tests/test_gpu_remap_edges.py tests/test_gpu_ragged.py tests/test_gpu_fused_batch.py tests/test_gpu_numpy_path.py tests/test_save_warped_image.py -m gpu -q -x > gpurun_out/r03f_pytest_map$pol.log 2>&1; echo "pytest exit $?" >> gpurun_out/r03f_pytest_map$pol.log
tail -8 gpurun_out/r03f_pytest_map$pol.log
done
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_remap_edges.py tests/test_gpu_ragged.py -m gpu -q -x -k "alignment or width_class or degenerate or odd or unsorted" > gpurun_out/r03f_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r03f_memcheck.log
tail -4 gpurun_out/r03f_memcheck.log
timeout 300 python profiles/s5_probe.py --only c --reps 40 2>&1 | tee gpurun_out/r03f_probe.txt
(timeout 300 python profiles/c4_probe.py; timeout 300 python profiles/c4_probe.py --round 4; timeout 300 python profiles/c4_probe.py --round 4 --odd) 2>&1 | tee gpurun_out/r03f_c4_probe.txt
