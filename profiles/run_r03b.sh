# round 2: MODE 2 reloads the row entry in place too; producer sleeping between failed polls (experiment)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_remap_edges.py tests/test_gpu_ragged.py tests/test_gpu_fused_batch.py -m gpu -q -x > gpurun_out/r03b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r03b_pytest.log
tail -4 gpurun_out/r03b_pytest.log
for ns in 0 50 200 500; do echo "== producer sleep $ns"; ATTWARP_QUAD_PRODUCER_SLEEP=$ns timeout 300 python profiles/s5_probe.py --only c --reps 40 2>&1 | tee gpurun_out/r03b_probe_sleep$ns.txt; done
timeout 300 python profiles/c4_probe.py 2>&1 | tee gpurun_out/r03b_c4_r1.txt
ATTWARP_QUAD_PRODUCER_SLEEP=200 timeout 300 python profiles/c4_probe.py 2>&1 | tee gpurun_out/r03b_c4_r1_sleep200.txt
