cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ATTWARP_FUZZ_CASES=1200 timeout 2400 python -m pytest tests/test_gpu_remap_fuzz.py -m gpu -q -n 6 -k "not other_formats" > gpurun_out/r04f_fuzz.log 2>&1; echo "pytest exit $?" >> gpurun_out/r04f_fuzz.log; tail -3 gpurun_out/r04f_fuzz.log | cut -c1-300
ATTWARP_FUZZ_CASES=160 timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_remap_fuzz.py -m gpu -q -x -k "uniform or ragged" > gpurun_out/r04f_fuzz_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r04f_fuzz_memcheck.log; tail -3 gpurun_out/r04f_fuzz_memcheck.log | cut -c1-300
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r04f_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r04f_pytest.log; tail -3 gpurun_out/r04f_pytest.log
timeout 900 python bench.py > gpurun_out/r04f_bench_default.json 2> /dev/null
python - <<'PY'
import json
d = json.load(open("gpurun_out/r04f_bench_default.json"))
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "sustained", round(d["sustained"]["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), d["roofline"]["kernel"], round(d["roofline"]["frac"], 3), {k: (round(v["ms"] * 1e3, 1), round(v["frac"], 3)) for k, v in d["kernels"].items()}, "c3", round(d["workloads"]["c3"]["value"]), round(d["workloads"]["c3"]["roofline"]["frac"], 3), "c4", round(d["workloads"]["c4"]["value"]), round(d["workloads"]["c4"]["roofline"]["frac"], 3))
PY
