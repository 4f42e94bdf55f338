# round 2: DIRECT stores at any row alignment (MODE 2) + the fused mask -> marginals kernel
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_remap_edges.py tests/test_gpu_ragged.py tests/test_gpu_fused_batch.py tests/test_gpu_numpy_path.py tests/test_mask_path.py tests/test_save_warped_image.py -m gpu -q -x > gpurun_out/r02r_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02r_pytest.log
tail -15 gpurun_out/r02r_pytest.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_remap_edges.py tests/test_gpu_ragged.py tests/test_mask_path.py -m gpu -q -x -k "alignment or width_class or degenerate or odd or fused" > gpurun_out/r02r_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r02r_memcheck.log
tail -4 gpurun_out/r02r_memcheck.log
echo "== default"; timeout 300 python profiles/s5_probe.py --only c --reps 40 2>&1 | tee gpurun_out/r02r_probe.txt
echo "== c4 round 1"; timeout 300 python profiles/c4_probe.py 2>&1 | tee gpurun_out/r02r_c4_r1.txt
echo "== c4 round 4"; timeout 300 python profiles/c4_probe.py --round 4 2>&1 | tee gpurun_out/r02r_c4_r4.txt
echo "== c4 round 1 tiles"; ATTWARP_QUAD_DIRECT=0 timeout 300 python profiles/c4_probe.py 2>&1 | tee gpurun_out/r02r_c4_r1_tiles.txt
