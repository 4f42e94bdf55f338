# two-GPU check of the sharded path (one process per GPU, NCCL only for timings/checksums)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for w in c2 c3 c4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload $w --no-cpu-baseline > gpurun_out/bench_${w}_2gpu.json 2> gpurun_out/bench_${w}_2gpu.err; tail -c 600 gpurun_out/bench_${w}_2gpu.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_2gpu.json 2> gpurun_out/bench_ref_2gpu.err
cat gpurun_out/bench_c2_2gpu.json gpurun_out/bench_c3_2gpu.json gpurun_out/bench_c4_2gpu.json gpurun_out/bench_ref_2gpu.json
