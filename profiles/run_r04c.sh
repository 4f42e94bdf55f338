cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ATTWARP_FUZZ_CASES=300 timeout 2400 python -m pytest tests/test_gpu_remap_fuzz.py tests/test_gpu_remap_edges.py tests/test_gpu_ragged.py tests/test_gpu_fused_batch.py tests/test_gpu_numpy_path.py -m gpu -q -n 6 > gpurun_out/r04c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r04c_pytest.log; tail -3 gpurun_out/r04c_pytest.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_remap_edges.py tests/test_gpu_ragged.py -m gpu -q -x -k "alignment or width_class or degenerate or odd or unsorted or minification" > gpurun_out/r04c_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r04c_memcheck.log; tail -2 gpurun_out/r04c_memcheck.log
timeout 300 python profiles/s5_probe.py --only c --reps 40 2>&1 | tee gpurun_out/r04c_probe.txt
(timeout 300 python profiles/c4_probe.py | tail -1; timeout 300 python profiles/c4_probe.py --round 4 | tail -1) 2>&1 | tee gpurun_out/r04c_c4.txt
timeout 600 python bench.py --workload c2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c2', round(d['value']), round(d['ms_per_step'],4), 'sustained', round(d['sustained']['ms_per_step'],4), {k:(round(v['ms']*1e3,1), round(v['frac'],3)) for k,v in d['kernels'].items()})"
