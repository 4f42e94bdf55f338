# round 2, first pass: the new parity tests, the new bench line (c2 headline + c3 + c4), the reference arm on the
# unmodified reference, stage-5 probe with the kernel's debug switches
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,temperature.gpu --format=csv > gpurun_out/r02a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02a_pytest_gpu.log
tail -5 gpurun_out/r02a_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02a_smoke.log 2>&1; tail -2 gpurun_out/r02a_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; tail -c 600 gpurun_out/r02a_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02a_bench_ref.json 2> gpurun_out/r02a_bench_ref.err; tail -c 300 gpurun_out/r02a_bench_ref.err
timeout 600 python profiles/s5_probe.py --dbg > gpurun_out/r02a_s5_probe.txt 2>&1
cat gpurun_out/r02a_s5_probe.txt
head -c 3000 gpurun_out/r02a_bench.json
