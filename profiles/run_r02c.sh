# round 2, third pass: quad kernel with prefetched row entries -- geometry / rows-per-chunk sweep
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_remap_edges.py tests/test_gpu_ragged.py tests/test_gpu_fused_batch.py -m gpu -q -x > gpurun_out/r02c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c_pytest.log
tail -4 gpurun_out/r02c_pytest.log
run() { echo "== $1" >> gpurun_out/r02c_probe.txt; env $1 timeout 300 python profiles/s5_probe.py --only $2 >> gpurun_out/r02c_probe.txt 2>&1; }
run "X=0" c2
run "ATTWARP_QUAD_ROWS=8" c2
run "ATTWARP_QUAD_ROWS=10" c2
run "ATTWARP_QUAD_GEO=4" c2
run "ATTWARP_QUAD_GEO=4 ATTWARP_QUAD_ROWS=7" c2
run "X=0" c3
run "ATTWARP_QUAD_GEO=2 ATTWARP_QUAD_ROWS=8" c3
run "ATTWARP_QUAD_GEO=3" c3
run "ATTWARP_QUAD_GEO=5" c3
run "ATTWARP_QUAD_GEO=1" c3
cat gpurun_out/r02c_probe.txt
