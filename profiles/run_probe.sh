cd $GRAFT_REPO_ROOT
python bench.py --no-cpu-baseline --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['steps'], d['clocks'])"
python bench.py --no-cpu-baseline --no-e2e --steps 20 --warmup 3 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['steps'], d['clocks'])"
python bench.py --workload c4 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['steps'], d['clocks'])"
