# producer warps sleep between failed polls: run time and executed instructions
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ATTWARP_FUZZ_CASES=200 timeout 2400 python -m pytest tests/test_gpu_remap_fuzz.py tests/test_gpu_remap_edges.py tests/test_gpu_remap_f32.py tests/test_gpu_ragged.py tests/test_gpu_fused_batch.py -m gpu -q -n 6 > gpurun_out/r04a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r04a_pytest.log; tail -3 gpurun_out/r04a_pytest.log
timeout 300 python profiles/s5_probe.py --only c --reps 40 2>&1 | tee gpurun_out/r04a_probe.txt
timeout 300 python bench.py --workload c5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c5', round(d['value']), {k:(round(v['ms']*1e3,1), round(v['frac'],3)) for k,v in d['kernels'].items()})"
timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:'remap_u8|remap_f32' -s 4 -c 2 --csv --log-file gpurun_out/r04a_inst.csv python profiles/s5_probe.py --only "c2  256x336^2 hwc near" --reps 4 > /dev/null 2>&1
grep "remap" gpurun_out/r04a_inst.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | cut -c1-200
