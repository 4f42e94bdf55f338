cd $GRAFT_REPO_ROOT
for w in c3 c5; do
python bench.py --workload $w --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w', round(d['value']), 'img/s', round(d['ms_per_step']*1e3,1), 'us/step', d['config']['launch'][:90])"
done
python bench.py --workload c5 --no-cpu-baseline --streams 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c5 1 stream', round(d['value']), 'img/s', round(d['ms_per_step']*1e3,1), 'us/step')"
