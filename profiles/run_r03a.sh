# round 2 (after pass z): next row's table entry loaded over the dead one (MODE 1); ragged class launches fanned out over 3 streams
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_remap_edges.py tests/test_gpu_ragged.py tests/test_gpu_fused_batch.py tests/test_gpu_numpy_path.py tests/test_image_io.py -m gpu -q -x > gpurun_out/r03a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r03a_pytest.log
tail -4 gpurun_out/r03a_pytest.log
timeout 300 python profiles/s5_probe.py --only c --reps 40 2>&1 | tee gpurun_out/r03a_probe.txt
for n in 3 2 1; do echo "== ragged streams $n"; ATTWARP_RAGGED_STREAMS=$n timeout 300 python profiles/c4_probe.py 2>&1 | tee gpurun_out/r03a_c4_r1_s$n.txt; ATTWARP_RAGGED_STREAMS=$n timeout 300 python profiles/c4_probe.py --round 4 2>&1 | tee gpurun_out/r03a_c4_r4_s$n.txt; done
timeout 600 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/r03a_bench_c4.json 2> gpurun_out/r03a_bench_c4.err; tail -c 300 gpurun_out/r03a_bench_c4.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r03a_bench_c4.json"))
print("c4 value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "roofline", round(d["roofline"]["frac"], 3))
PY
