# round 2, last pass r09z: the driver's own commands on the final tree
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
P=gpurun_out/r09z
timeout 1800 python -m pytest tests/ -x -q -m gpu > ${P}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> ${P}_pytest_gpu.log; tail -n 4 ${P}_pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1; tail -n 1 ${P}_smoke.log
timeout 900 python bench.py > ${P}_bench_default.json 2> ${P}_bench_default.err; tail -c 200 ${P}_bench_default.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r09z_bench_default.json"))
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), d["roofline"]["kernel"], round(d["roofline"]["frac"], 3),
      {k: (round(v["ms"] * 1e3, 1), round(v["frac"], 3)) for k, v in d["kernels"].items()},
      "c3", round(d["workloads"]["c3"]["value"]), round(d["workloads"]["c3"]["roofline"]["frac"], 3),
      "c4", round(d["workloads"]["c4"]["value"]), round(d["workloads"]["c4"]["roofline"]["frac"], 3),
      "flow", round(d["driver_flow"]["step_ms"] * 1e3, 1), round(d["workloads"]["c3"]["driver_flow"]["step_ms"] * 1e3, 1), "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"],
      "cpu", round(d["cpu_baseline"]["value"]), d["cpu_baseline"]["kind"])
PY
