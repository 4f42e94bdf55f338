# round 2: DIRECT launches with exactly the consumer warps a strip needs (3..16), ragged classes per 128 columns
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_remap_edges.py tests/test_gpu_ragged.py tests/test_gpu_fused_batch.py tests/test_gpu_numpy_path.py -m gpu -q -x > gpurun_out/r02q_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02q_pytest.log
tail -5 gpurun_out/r02q_pytest.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_ragged.py -m gpu -q -x -k "width_class or degenerate" > gpurun_out/r02q_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r02q_memcheck.log
tail -4 gpurun_out/r02q_memcheck.log
echo "== default"; timeout 300 python profiles/s5_probe.py --reps 40 2>&1 | tee gpurun_out/r02q_probe.txt
for px in 0 2000000 8000000; do
echo "== c4 round 4, class min px $px"; ATTWARP_QUAD_CLASS_MIN_PX=$px timeout 300 python profiles/c4_probe.py --round 4 2>&1 | tee gpurun_out/r02q_c4_r4_px$px.txt
done
echo "== c4 round 4, max 11 warps"; ATTWARP_QUAD_MAXW=11 timeout 300 python profiles/c4_probe.py --round 4 2>&1 | tee gpurun_out/r02q_c4_r4_maxw11.txt
echo "== c4 round 1"; timeout 300 python profiles/c4_probe.py 2>&1 | tee gpurun_out/r02q_c4_r1.txt
echo "== c4 round 4 1024..2048"; timeout 300 python profiles/c4_probe.py --round 4 --min-side 1409 --n 256 2>&1 | tee gpurun_out/r02q_c4_r4_wide.txt
echo "== c4 round 4 705..1408"; timeout 300 python profiles/c4_probe.py --round 4 --min-side 705 --max-side 1408 --n 512 2>&1 | tee gpurun_out/r02q_c4_r4_mid.txt
