# round 2: aligned classes join the any-alignment class of the same width; <384,1> build for 10..11 warps
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_remap_edges.py tests/test_gpu_ragged.py -m gpu -q -x > gpurun_out/r02s_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02s_pytest.log
tail -5 gpurun_out/r02s_pytest.log
echo "== default"; timeout 300 python profiles/s5_probe.py --only c --reps 40 2>&1 | tee gpurun_out/r02s_probe.txt
echo "== c4 round 1"; timeout 300 python profiles/c4_probe.py 2>&1 | tee gpurun_out/r02s_c4_r1.txt
echo "== c4 round 1 no join"; ATTWARP_QUAD_ALIGNED_MIN_PX=0 timeout 300 python profiles/c4_probe.py 2>&1 | tee gpurun_out/r02s_c4_r1_nojoin.txt
echo "== c4 round 4"; timeout 300 python profiles/c4_probe.py --round 4 2>&1 | tee gpurun_out/r02s_c4_r4.txt
echo "== c4 odd sides only (all unaligned)"; timeout 300 python profiles/c4_probe.py --round 4 --odd 2>&1 | tee gpurun_out/r02s_c4_odd.txt
