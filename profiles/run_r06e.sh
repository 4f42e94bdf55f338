# round 2, pass r06e: walk kernel walks all planes of a planar image per thread; grey images back on the round-1 kernel; square back to float64
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
P=gpurun_out/r06e
timeout 900 python -m pytest tests/test_gpu_remap_edges.py tests/test_gpu_numpy_path.py tests/test_gpu_torch_path.py -m gpu -q -x -n 4 > ${P}_pytest_new.log 2>&1; echo "pytest exit $?" >> ${P}_pytest_new.log; tail -n 5 ${P}_pytest_new.log | cut -c1-400
ATTWARP_FUZZ_CASES=400 timeout 900 python -m pytest tests/test_gpu_remap_fuzz.py -m gpu -q -n 6 -k "other_formats" > ${P}_fuzz_formats.log 2>&1; echo "pytest exit $?" >> ${P}_fuzz_formats.log; tail -n 3 ${P}_fuzz_formats.log | cut -c1-400
timeout 600 python profiles/formats_probe.py 2>&1 | grep -v Warning > ${P}_formats_probe.txt; cat ${P}_formats_probe.txt
timeout 600 python profiles/row_kernels_probe.py 2>&1 | grep -v Warning | head -3 > ${P}_row_kernels.txt; cat ${P}_row_kernels.txt
