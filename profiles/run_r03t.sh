cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ATTWARP_FUZZ_CASES=1600 timeout 2400 python -m pytest tests/test_gpu_remap_fuzz.py -m gpu -q -n 6 > gpurun_out/r03t_fuzz.log 2>&1; echo "pytest exit $?" >> gpurun_out/r03t_fuzz.log; tail -6 gpurun_out/r03t_fuzz.log | cut -c1-300
ATTWARP_REMAP_F32=rows ATTWARP_FUZZ_CASES=800 timeout 2400 python -m pytest tests/test_gpu_remap_fuzz.py -m gpu -q -n 6 -k other_formats > gpurun_out/r03t_fuzz_rows.log 2>&1; echo "pytest exit $?" >> gpurun_out/r03t_fuzz_rows.log; tail -4 gpurun_out/r03t_fuzz_rows.log | cut -c1-300
ATTWARP_REMAP_QUAD=0 ATTWARP_FUZZ_CASES=400 timeout 2400 python -m pytest tests/test_gpu_remap_fuzz.py -m gpu -q -n 6 -k "uniform or ragged" > gpurun_out/r03t_fuzz_stream.log 2>&1; echo "pytest exit $?" >> gpurun_out/r03t_fuzz_stream.log; tail -4 gpurun_out/r03t_fuzz_stream.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_remap_f32.py tests/test_c5_marginalnet.py tests/test_gpu_torch_path.py -m gpu -q > gpurun_out/r03t_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r03t_pytest.log; tail -3 gpurun_out/r03t_pytest.log
