# round 2: float32 streaming resample kernel -- parity, then timing vs the round-1 rows kernel
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_remap_f32.py tests/test_gpu_torch_path.py tests/test_c5_marginalnet.py tests/test_gpu_fused_batch.py -m gpu -q -x > gpurun_out/r02o_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02o_pytest.log
tail -30 gpurun_out/r02o_pytest.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_remap_f32.py -m gpu -q -x -k "unsorted or degenerate or odd" > gpurun_out/r02o_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r02o_memcheck.log
tail -4 gpurun_out/r02o_memcheck.log
timeout 300 python bench.py --workload c5 --no-cpu-baseline > gpurun_out/r02o_bench_c5.json 2> gpurun_out/r02o_bench_c5.err; tail -c 300 gpurun_out/r02o_bench_c5.err
ATTWARP_REMAP_F32=rows timeout 300 python bench.py --workload c5 --no-cpu-baseline > gpurun_out/r02o_bench_c5_rows.json 2>/dev/null
for r in 8 12; do ATTWARP_F32_ROWS=$r timeout 300 python bench.py --workload c5 --no-cpu-baseline > gpurun_out/r02o_bench_c5_r$r.json 2>/dev/null; done
python - <<'PY'
import json
for f in ("r02o_bench_c5", "r02o_bench_c5_rows", "r02o_bench_c5_r8", "r02o_bench_c5_r12"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value", round(d["value"]), "ms", round(d["ms_per_step"], 4), {k: (round(v["ms"] * 1e3, 1), round(v["frac"], 3)) for k, v in d["kernels"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
