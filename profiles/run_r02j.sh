# round 2: pipelined soft store path, per-wave skew
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_remap_edges.py tests/test_gpu_ragged.py tests/test_gpu_fused_batch.py -m gpu -q -x > gpurun_out/r02j_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02j_pytest.log
tail -4 gpurun_out/r02j_pytest.log
run() { echo "== $1" >> gpurun_out/r02j_probe.txt; env $1 timeout 300 python profiles/s5_probe.py --only "$2" >> gpurun_out/r02j_probe.txt 2>&1; }
run "ATTWARP_QUAD_SKEW_PPM=0" "c2  256x336^2 hwc near"
run "ATTWARP_QUAD_SKEW_PPM=40000" "c2  256x336^2 hwc near"
run "ATTWARP_QUAD_SKEW_PPM=60000" "c"
run "ATTWARP_QUAD_SKEW_PPM=80000" "c2  256x336^2 hwc near"
run "ATTWARP_QUAD_SKEW_PPM=110000" "c2  256x336^2 hwc near"
run "ATTWARP_QUAD_SKEW_PPM=60000" "    1024"
cat gpurun_out/r02j_probe.txt
timeout 300 python profiles/c4_probe.py > gpurun_out/r02j_c4.txt 2>&1; cat gpurun_out/r02j_c4.txt
