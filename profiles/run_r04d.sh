cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ATTWARP_FUZZ_CASES=1600 timeout 2400 python -m pytest tests/test_gpu_remap_fuzz.py -m gpu -q -n 6 -k "not other_formats" > gpurun_out/r04d_fuzz.log 2>&1; echo "pytest exit $?" >> gpurun_out/r04d_fuzz.log; tail -4 gpurun_out/r04d_fuzz.log | cut -c1-300
for pol in 1 2; do ATTWARP_QUAD_MAP=$pol ATTWARP_FUZZ_CASES=400 timeout 1200 python -m pytest tests/test_gpu_remap_fuzz.py -m gpu -q -n 6 -k "uniform or ragged" > gpurun_out/r04d_fuzz_map$pol.log 2>&1; echo "pytest exit $?" >> gpurun_out/r04d_fuzz_map$pol.log; tail -2 gpurun_out/r04d_fuzz_map$pol.log | cut -c1-300; done
ATTWARP_FUZZ_CASES=240 timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_remap_fuzz.py -m gpu -q -x -k "uniform or ragged" > gpurun_out/r04d_fuzz_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r04d_fuzz_memcheck.log; tail -3 gpurun_out/r04d_fuzz_memcheck.log | cut -c1-300
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r04d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r04d_pytest.log; tail -3 gpurun_out/r04d_pytest.log
