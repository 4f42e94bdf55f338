# round 2: output tiles + store warps removed from the uint8 kernel (MODE 1 / MODE 2 only): full GPU suite, memcheck, bench
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r02t_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02t_pytest.log
tail -6 gpurun_out/r02t_pytest.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_remap_edges.py tests/test_gpu_ragged.py tests/test_mask_path.py -m gpu -q -x -k "alignment or width_class or degenerate or odd or fused or unsorted or minification" > gpurun_out/r02t_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r02t_memcheck.log
tail -4 gpurun_out/r02t_memcheck.log
timeout 300 python profiles/s5_probe.py --reps 40 --dbg 2>&1 | tee gpurun_out/r02t_probe.txt
timeout 300 python profiles/c4_probe.py 2>&1 | tee gpurun_out/r02t_c4_r1.txt
timeout 300 python profiles/c4_probe.py --round 4 2>&1 | tee gpurun_out/r02t_c4_r4.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02t_bench.json 2> gpurun_out/r02t_bench.err; tail -c 600 gpurun_out/r02t_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02t_bench.json"))
print("value", round(d["value"]), d["unit"], "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]))
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac") if k in d["roofline"]})
print("kernels", {k: (round(v["ms"] * 1e3, 1), round(v["frac"], 3)) for k, v in d.get("kernels", {}).items()})
for k, w in d.get("workloads", {}).items():
    print(k, "value", round(w["value"]), "ms", round(w.get("ms_per_step", 0), 4), "roofline", round(w["roofline"]["frac"], 3), {kk: (round(v["ms"] * 1e3, 1), round(v["frac"], 3)) for kk, v in w.get("kernels", {}).items()})
PY
