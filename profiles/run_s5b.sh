cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s5_pytest.log 2>&1; tail -3 gpurun_out/s5_pytest.log
for w in c2 c3 c5; do
timeout 600 python bench.py --workload $w --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -c 600 gpurun_out/bench_$w.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$w.json"))
print("$w", round(d["value"]), "img/s", round(d["ms_per_step"]*1e3,1), "us/step", {k:(round(v["ms"]*1e3,1), round(v["frac"],3)) for k,v in d.get("kernels",{}).items()}, d["roofline"]["kernel"], round(d["roofline"]["frac"],3), d["e2e"] and round(d["e2e"]["value"]))
PY
done
