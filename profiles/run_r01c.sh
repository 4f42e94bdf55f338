set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload c2 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 600 gpurun_out/bench_c2.err
timeout 600 python bench.py --workload c3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-graph > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:remap_u8_tiled -s 2 -c 1 -o gpurun_out/prof_remap336 -f python profiles/drive.py remap --side 336 --batch 256 > gpurun_out/ncu_remap336.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:remap_u8_tiled -s 2 -c 1 -o gpurun_out/prof_remap1344 -f python profiles/drive.py remap --side 1344 --batch 64 > gpurun_out/ncu_remap1344.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:aggregate -s 2 -c 1 -o gpurun_out/prof_agg -f python profiles/drive.py agg > gpurun_out/ncu_agg.log 2>&1
python profiles/drive.py remap --side 336 --batch 256 --iters 10
python profiles/drive.py remap --side 1344 --batch 64 --iters 10
python profiles/drive.py remap --side 512 --batch 128 --dtype f32 --layout chw --iters 5
python profiles/drive.py agg --iters 10
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
cat gpurun_out/bench_c2.json
