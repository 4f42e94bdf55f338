# round-1 (fifth session, final state) measurement pass: GPU tests, bench lines for every workload,
# ncu launch list and full captures of the dominant kernels, secondary-kernel survey
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --workload c2 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 400 gpurun_out/bench_c2.err
timeout 600 python bench.py --workload c2 --streams 1 --rotate 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c2_1stream.json 2> gpurun_out/bench_c2_1stream.err
for w in c3 c4 c5; do
timeout 600 python bench.py --workload $w --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-graph --streams 1 --rotate 3 > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'aggregate_rows|maps_from|remap_u8' -s 6 -c 3 -o gpurun_out/prof_c2 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-graph --streams 1 --rotate 3 > gpurun_out/ncu_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:remap_u8 -s 2 -c 1 -o gpurun_out/prof_remap1344 -f python profiles/drive.py remap --side 1344 --batch 64 > gpurun_out/ncu_remap1344.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:remap_f32 -s 2 -c 1 -o gpurun_out/prof_remapf32 -f python profiles/drive.py remap --side 512 --batch 128 --dtype f32 --layout chw > gpurun_out/ncu_remapf32.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'marginals_u8|maps_from_partials' -s 2 -c 2 -o gpurun_out/prof_att1344 -f python profiles/drive.py att --side 1344 --batch 64 > gpurun_out/ncu_att.log 2>&1
timeout 600 python profiles/kernel_survey.py > gpurun_out/kernel_survey.txt 2>&1
cat gpurun_out/bench_c2.json gpurun_out/bench_c2_1stream.json gpurun_out/bench_c3.json gpurun_out/bench_c4.json gpurun_out/bench_c5.json gpurun_out/bench_ref.json
