cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r03x_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r03x_pytest.log; tail -4 gpurun_out/r03x_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r03x_bench_default.json 2> gpurun_out/r03x_bench_default.err; tail -c 200 gpurun_out/r03x_bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r03x_bench_ref.json 2>/dev/null
python - <<'PY'
import json
d = json.load(open("gpurun_out/r03x_bench_default.json"))
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), d["roofline"]["kernel"], round(d["roofline"]["frac"], 3), {k: (round(v["ms"] * 1e3, 1), round(v["frac"], 3)) for k, v in d["kernels"].items()}, "c3", round(d["workloads"]["c3"]["value"]), "c4", round(d["workloads"]["c4"]["value"]), "launches", d["gpu_launches"])
r = json.load(open("gpurun_out/r03x_bench_ref.json")); print("reference", round(r["value"]), r["cpu_baseline"]["kind"])
PY
