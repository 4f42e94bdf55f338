# round 2: quad kernel ring-depth experiments + new tests (PDF-L1 loss, RaggedBatch)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pdf_loss.py tests/test_gpu_ragged.py tests/test_gpu_remap_edges.py tests/test_gpu_fused_batch.py -m gpu -q -x > gpurun_out/r02e_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02e_pytest.log
tail -12 gpurun_out/r02e_pytest.log
run() { echo "== $1" >> gpurun_out/r02e_probe.txt; env $1 timeout 300 python profiles/s5_probe.py --only "$2" >> gpurun_out/r02e_probe.txt 2>&1; }
run "X=0" ""
run "ATTWARP_QUAD_RING=33" c3
run "ATTWARP_QUAD_RING=32" c3
run "ATTWARP_QUAD_RING=23" c3
run "ATTWARP_QUAD_RING=44" c3
run "ATTWARP_QUAD_RING=33" c2
run "ATTWARP_QUAD_RING=32" c2
run "ATTWARP_QUAD_RING=33 ATTWARP_QUAD_GEO=4" c2
cat gpurun_out/r02e_probe.txt
