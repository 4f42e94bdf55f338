# round 2, second pass: the new 4-pixels-per-thread stage-5 kernel (remap_quad.cu): parity tests, memcheck on the
# edge suite, timing probe per geometry
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_remap_edges.py tests/test_gpu_numpy_path.py tests/test_gpu_ragged.py tests/test_gpu_fused_batch.py tests/test_save_warped_image.py tests/test_mask_path.py tests/test_gpu_torch_path.py -m gpu -q -x > gpurun_out/r02b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02b_pytest.log
tail -25 gpurun_out/r02b_pytest.log
timeout 600 python profiles/s5_probe.py --dbg > gpurun_out/r02b_probe_quad.txt 2>&1; cat gpurun_out/r02b_probe_quad.txt
for g in 0 1; do
ATTWARP_QUAD_GEO=$g timeout 300 python profiles/s5_probe.py --only c3 > gpurun_out/r02b_probe_geo$g.txt 2>&1; cat gpurun_out/r02b_probe_geo$g.txt
done
for r in 8 12; do
ATTWARP_QUAD_ROWS=$r timeout 300 python profiles/s5_probe.py --only c2 > gpurun_out/r02b_probe_rows$r.txt 2>&1; cat gpurun_out/r02b_probe_rows$r.txt
done
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_remap_edges.py -m gpu -q -x > gpurun_out/r02b_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r02b_memcheck.log
tail -15 gpurun_out/r02b_memcheck.log
