# round 2: co-residency of stage 1 and stage 5 in the 4-stream step (shared memory per SM is the limit)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { echo "== $1 | $2" >> gpurun_out/r02n_overlap.txt; env $1 timeout 300 python bench.py --workload c2 --no-cpu-baseline --no-e2e $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', round(d['value']), 'ms/step', round(d['ms_per_step']*1e3,1), 'sustained', round(d['sustained']['ms_per_step']*1e3,1), 'single', round(d['single_stream']['ms_per_step']*1e3,1), {k: round(v['ms']*1e3,1) for k,v in d['kernels'].items()})
" >> gpurun_out/r02n_overlap.txt 2>&1; }
run "X=0" ""
run "ATTWARP_REMAP_CTAS_PER_SM=2" ""
run "ATTWARP_REMAP_CTAS_PER_SM=3" ""
run "ATTWARP_AGG_NSPLIT=1" ""
run "ATTWARP_AGG_NSPLIT=1 ATTWARP_REMAP_CTAS_PER_SM=2" ""
run "ATTWARP_AGG_NSPLIT=1 ATTWARP_REMAP_CTAS_PER_SM=3" ""
run "X=0" "--streams 2"
run "X=0" "--streams 3"
run "X=0" "--streams 6 --rotate 12"
run "ATTWARP_AGG_NSPLIT=1 ATTWARP_REMAP_CTAS_PER_SM=2" "--streams 6 --rotate 12"
cat gpurun_out/r02n_overlap.txt
