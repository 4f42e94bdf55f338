# round 2, pass r08b: float32 marginals: max.NaN clamp, no predicate on zero-filled pixels, ping-pong of the rows in flight
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
P=gpurun_out/r08b
timeout 900 python -m pytest tests/test_gpu_numpy_path.py tests/test_gpu_torch_path.py tests/test_gpu_stage_fuzz.py tests/test_gpu_autograd.py tests/test_c5_marginalnet.py -m gpu -q -n 4 > ${P}_pytest.log 2>&1; echo "pytest exit $?" >> ${P}_pytest.log; tail -n 6 ${P}_pytest.log | cut -c1-500
timeout 600 python profiles/formats_probe.py 2>&1 | grep -v Warning | grep "maps_from_attention" > ${P}_formats_probe.txt; cat ${P}_formats_probe.txt
timeout 600 python profiles/row_kernels_probe.py 2>&1 | grep -v Warning | head -2 > ${P}_row_kernels.txt; cat ${P}_row_kernels.txt
