# round 2: horizontal blend in 6 instructions per pixel and slot (channel 0 straight from the aligned word)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ATTWARP_FUZZ_CASES=300 timeout 2400 python -m pytest tests/test_gpu_remap_fuzz.py tests/test_gpu_remap_edges.py tests/test_gpu_ragged.py tests/test_gpu_fused_batch.py -m gpu -q -n 6 -k "not other_formats" > gpurun_out/r03y_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r03y_pytest.log; tail -4 gpurun_out/r03y_pytest.log
timeout 300 python profiles/s5_probe.py --only c --reps 40 2>&1 | tee gpurun_out/r03y_probe.txt
(timeout 300 python profiles/c4_probe.py | tail -1; timeout 300 python profiles/c4_probe.py --round 4 | tail -1) 2>&1 | tee gpurun_out/r03y_c4.txt
