# round 2, pass r08c: bench line with the labelled driver-flow extra
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
P=gpurun_out/r08c
timeout 900 python bench.py --steps 20 --warmup 5 > ${P}_bench.json 2> ${P}_bench.err; tail -c 300 ${P}_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r08c_bench.json"))
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3))
print("c2 driver_flow", d.get("driver_flow"))
print("c3 driver_flow", d["workloads"]["c3"].get("driver_flow"))
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"])
PY
