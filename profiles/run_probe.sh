cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python bench.py --workload c2 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 300 gpurun_out/bench_c2.err
python -c "import json; d=json.load(open('gpurun_out/bench_c2.json')); print(round(d['value']), d['clocks'], round(d['e2e']['value']))"
timeout 600 python bench.py --workload c2 --steps 20 --warmup 3 --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['steps'], d['clocks'])"
