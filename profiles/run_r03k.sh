cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r03k_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r03k_pytest.log
tail -4 gpurun_out/r03k_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r03k_bench_default.json 2> gpurun_out/r03k_bench_default.err; tail -c 300 gpurun_out/r03k_bench_default.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r03k_bench_default.json"))
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), {k: (round(v["ms"] * 1e3, 1), round(v["frac"], 3)) for k, v in d["kernels"].items()}, "c3", round(d["workloads"]["c3"]["value"]), "c4", round(d["workloads"]["c4"]["value"]), "launches", d["gpu_launches"])
PY
