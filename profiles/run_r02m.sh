# round 2, measurement pass m: all GPU tests, smoke, the bench line (c2 + c3 + c4), reference arm, c5, survey,
# ncu launch list and full captures of the step's kernels
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02m_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02m_pytest_gpu.log
tail -6 gpurun_out/r02m_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02m_smoke.log 2>&1; tail -1 gpurun_out/r02m_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err; tail -c 400 gpurun_out/r02m_bench.err
timeout 600 python bench.py > gpurun_out/r02m_bench_default.json 2> gpurun_out/r02m_bench_default.err; tail -c 400 gpurun_out/r02m_bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02m_bench_ref.json 2> gpurun_out/r02m_bench_ref.err
timeout 600 python bench.py --workload c5 --no-cpu-baseline > gpurun_out/r02m_bench_c5.json 2> gpurun_out/r02m_bench_c5.err
timeout 600 python profiles/kernel_survey.py > gpurun_out/r02m_kernel_survey.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02m_launches_c2.csv python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-graph --streams 1 --rotate 3 > gpurun_out/r02m_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'aggregate_rows|maps_from|remap_u8' -s 6 -c 3 -o gpurun_out/r02m_prof_c2 -f python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-graph --streams 1 --rotate 3 > gpurun_out/r02m_ncu_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:remap_u8 -s 2 -c 1 -o gpurun_out/r02m_prof_remap1344 -f python profiles/drive.py remap --side 1344 --batch 64 --grid 48 > gpurun_out/r02m_ncu_remap1344.log 2>&1
python - <<'PY'
import json
for f in ("r02m_bench", "r02m_bench_default"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), {k: round(v["frac"], 3) for k, v in d["kernels"].items()},
              "c3", round(d["workloads"]["c3"]["value"]), round(d["workloads"]["c3"]["roofline"]["frac"], 3), "c4", round(d["workloads"]["c4"]["value"]), round(d["workloads"]["c4"]["roofline"]["frac"], 3))
    except Exception as e:
        print(f, "failed", e)
PY
cat gpurun_out/r02m_kernel_survey.txt
