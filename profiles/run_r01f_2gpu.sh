# two-GPU check of the final state (one process per GPU under torchrun, NCCL only for timings/checksums)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for w in c2 c3; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload $w --no-cpu-baseline > gpurun_out/bench_${w}_2gpu.json 2> gpurun_out/bench_${w}_2gpu.err; wc -l gpurun_out/bench_${w}_2gpu.json; tail -c 300 gpurun_out/bench_${w}_2gpu.err | grep -v "OMP_NUM\|^\*\*\*" 
python -c "import json; d=json.load(open('gpurun_out/bench_${w}_2gpu.json')); print('$w', round(d['value']), 'img/s', d['n_gpus'], 'gpus', round(d['ms_per_step']*1e3,1), 'us/step', d['clocks'] and d['clocks']['samples'], d.get('e2e') and round(d['e2e']['value']))"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_2gpu.json 2> gpurun_out/bench_ref_2gpu.err
python -c "import json; d=json.load(open('gpurun_out/bench_ref_2gpu.json')); print('ref', round(d['value']))"
for st in 4; do
python bench.py --workload c2 --no-cpu-baseline --no-e2e --streams $st --rotate 12 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c2 streams $st', round(d['value']), round(d['ms_per_step']*1e3,1))"
done
