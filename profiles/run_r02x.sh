# round 2: aligned rows stored by their own lanes (12-byte lane stride) in both mappings; LANE paths read their
# four adjacent RGBX pixels with one 128-bit load
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_remap_edges.py tests/test_gpu_ragged.py tests/test_gpu_fused_batch.py tests/test_gpu_numpy_path.py -m gpu -q -x > gpurun_out/r02x_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02x_pytest.log
tail -4 gpurun_out/r02x_pytest.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_remap_edges.py tests/test_gpu_ragged.py -m gpu -q -x -k "alignment or width_class or degenerate or odd" > gpurun_out/r02x_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r02x_memcheck.log
tail -3 gpurun_out/r02x_memcheck.log
for p in 0 1 2; do echo "== policy $p"; ATTWARP_QUAD_MAP=$p timeout 300 python profiles/s5_probe.py --only c --reps 40 2>&1 | tee gpurun_out/r02x_probe_map$p.txt; done
timeout 300 python profiles/c4_probe.py 2>&1 | tee gpurun_out/r02x_c4_r1.txt
timeout 300 python profiles/c4_probe.py --round 4 2>&1 | tee gpurun_out/r02x_c4_r4.txt
