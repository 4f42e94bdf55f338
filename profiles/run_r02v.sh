# round 2: scratch single-buffered (a warp barrier before the writes instead of xor-toggled double buffers)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_remap_edges.py tests/test_gpu_ragged.py tests/test_gpu_fused_batch.py tests/test_gpu_numpy_path.py -m gpu -q -x > gpurun_out/r02v_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02v_pytest.log
tail -4 gpurun_out/r02v_pytest.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_remap_edges.py -m gpu -q -x -k "alignment and 335 or mappings_agree and 97" > gpurun_out/r02v_racecheck.log 2>&1; echo "racecheck exit $?" >> gpurun_out/r02v_racecheck.log
tail -5 gpurun_out/r02v_racecheck.log
timeout 300 python profiles/s5_probe.py --only c --reps 40 2>&1 | tee gpurun_out/r02v_probe.txt
timeout 300 python profiles/c4_probe.py 2>&1 | tee gpurun_out/r02v_c4_r1.txt
timeout 300 python profiles/c4_probe.py --round 4 2>&1 | tee gpurun_out/r02v_c4_r4.txt
