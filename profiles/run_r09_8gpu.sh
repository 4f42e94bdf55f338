# 8-GPU run of the driver's bench command on the final tree
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
P=gpurun_out/r09_8gpu
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 > ${P}_bench.json 2> ${P}_bench.err; tail -c 300 ${P}_bench.err | grep -v "OMP_NUM\|^\*\*\*"
python - <<PY
import json
d = json.load(open("${P}_bench.json"))
print("N=8 c2 value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "frac_of_copy_ceiling", round(d["e2e"]["frac_of_copy_ceiling"], 3),
      "c3", round(d["workloads"]["c3"]["value"]), "c4", round(d["workloads"]["c4"]["value"]), [round(x, 3) for x in d["workloads"]["c4"]["per_rank_ms"]], d["workloads"]["c4"]["check"]["sharded_equals_unsharded"],
      "flow", round(d["driver_flow"]["step_ms"], 4), round(d["workloads"]["c3"]["driver_flow"]["step_ms"], 4), "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
