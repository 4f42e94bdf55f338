cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s5_pytest.log 2>&1; tail -3 gpurun_out/s5_pytest.log
timeout 600 python bench.py --workload c2 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 600 gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
timeout 600 python bench.py --workload c2 --no-cpu-baseline --no-e2e --streams 1 --rotate 3
