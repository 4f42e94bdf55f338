# round 2, measurement pass r07z (after the marginals and tensor-core LANCZOS work): all GPU tests, smoke, the bench line (c2 + c3 + c4), reference arm, c5, survey, probes,
# ncu launch lists and full captures of the step's kernels
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
P=gpurun_out/r07z
timeout 1500 python -m pytest tests -m gpu -q -n 4 > ${P}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> ${P}_pytest_gpu.log
tail -6 ${P}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > ${P}_smoke.log 2>&1; tail -1 ${P}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > ${P}_bench.json 2> ${P}_bench.err; tail -c 300 ${P}_bench.err
timeout 900 python bench.py > ${P}_bench_default.json 2> ${P}_bench_default.err; tail -c 300 ${P}_bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > ${P}_bench_ref.json 2> ${P}_bench_ref.err
timeout 600 python bench.py --workload c5 --no-cpu-baseline > ${P}_bench_c5.json 2> ${P}_bench_c5.err
timeout 600 python profiles/kernel_survey.py > ${P}_kernel_survey.txt 2>&1
timeout 300 python profiles/s5_probe.py --reps 40 --dbg > ${P}_s5_probe.txt 2>&1
(timeout 300 python profiles/c4_probe.py; timeout 300 python profiles/c4_probe.py --round 4; timeout 300 python profiles/c4_probe.py --round 4 --odd) > ${P}_c4_probe.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${P}_launches_c2.csv python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-graph --streams 1 --rotate 3 > ${P}_ncu_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'maps_from|remap_u8' -c 200 --csv --log-file ${P}_launches_c4.csv python profiles/drive.py ragged --batch 1024 --iters 2 > ${P}_ncu_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'aggregate_rows|maps_from|remap_u8' -s 6 -c 3 -o ${P}_prof_c2 -f python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-graph --streams 1 --rotate 3 > ${P}_ncu_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:remap_u8 -s 2 -c 1 -o ${P}_prof_remap1344 -f python profiles/drive.py remap --side 1344 --batch 64 --grid 48 > ${P}_ncu_remap1344.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:remap_f32 -s 2 -c 1 -o ${P}_prof_remapf32 -f python profiles/drive.py remap --side 512 --batch 128 --dtype f32 --layout chw > ${P}_ncu_remapf32.log 2>&1
python - <<'PY'
import json
for f in ("r07z_bench", "r07z_bench_default"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "sustained", round(d["sustained"]["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]),
              {k: (round(v["ms"] * 1e3, 1), round(v["frac"], 3)) for k, v in d["kernels"].items()},
              "c3", round(d["workloads"]["c3"]["value"]), round(d["workloads"]["c3"]["roofline"]["frac"], 3), "c4", round(d["workloads"]["c4"]["value"]), round(d["workloads"]["c4"]["roofline"]["frac"], 3))
    except Exception as e:
        print(f, "failed", e)
for f in ("r07z_bench_ref", "r07z_bench_c5"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value", round(d["value"]), d.get("cpu_baseline", {}).get("kind"), {k: (round(v["ms"] * 1e3, 1), round(v["frac"], 3)) for k, v in d.get("kernels", {}).items()})
    except Exception as e:
        print(f, "failed", e)
PY
cat ${P}_c4_probe.txt
timeout 600 python profiles/row_kernels_probe.py 2>&1 | grep -v Warning > ${P}_row_kernels.txt
timeout 600 python profiles/formats_probe.py 2>&1 | grep -v Warning > ${P}_formats_probe.txt
