cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ragged.py tests/test_image_io.py tests/test_sharding_gloo.py -m gpu -q -x > gpurun_out/r03n_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r03n_pytest.log; tail -3 gpurun_out/r03n_pytest.log
timeout 600 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/r03n_bench_c4.json 2> gpurun_out/r03n_bench_c4.err; tail -c 200 gpurun_out/r03n_bench_c4.err
python -c "import json; d=json.load(open('gpurun_out/r03n_bench_c4.json')); print('c4', round(d['value']), d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'])"
