# N-GPU record of the final state: the driver's own bench line (c2 + c3 + c4), the reference arm, the bare-copy ceiling
# usage: bash profiles/run_r03_ngpu.sh N
N=$1
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
P=gpurun_out/r05_${N}gpu
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > ${P}_bench.json 2> ${P}_bench.err; tail -c 300 ${P}_bench.err | grep -v "OMP_NUM\|^\*\*\*"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --impl reference --steps 2 --warmup 1 > ${P}_bench_ref.json 2> ${P}_bench_ref.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 profiles/h2d_ceiling.py > ${P}_h2d_ceiling.txt 2> ${P}_h2d_ceiling.err; tail -5 ${P}_h2d_ceiling.txt
python - <<PY
import json
d = json.load(open("${P}_bench.json"))
print("N=$N c2 value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "sustained", round(d["sustained"]["value"]), "e2e", round(d["e2e"]["value"]), "frac_of_copy_ceiling", round(d["e2e"]["frac_of_copy_ceiling"], 3),
      "c3", round(d["workloads"]["c3"]["value"]), "c3 e2e", round(d["workloads"]["c3"]["e2e"]["value"]), "c4", round(d["workloads"]["c4"]["value"]), d["workloads"]["c4"]["per_rank_ms"], d["workloads"]["c4"]["check"]["sharded_equals_unsharded"])
r = json.load(open("${P}_bench_ref.json"))
print("reference arm", round(r["value"]), r["cpu_baseline"]["cores"], "cores")
PY
