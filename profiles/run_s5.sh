# stage-5 iteration: parity tests that touch the uint8 resample kernel, event timings, light ncu counters
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s5_pytest.log 2>&1; tail -3 gpurun_out/s5_pytest.log
timeout 120 python profiles/drive.py remap --side 336 --batch 256 --iters 12 | sed 's/GB.*//'
timeout 120 python profiles/drive.py remap --side 1344 --batch 64 --iters 8 | sed 's/GB.*//'
timeout 120 python profiles/drive.py ragged --batch 128 --iters 4
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
timeout 300 ncu --metrics $M --clock-control none -k regex:remap_u8 -s 2 -c 1 python profiles/drive.py remap --side 336 --batch 256 2>&1 | grep -E "gpu__time|smsp__|l1tex" 
timeout 300 ncu --metrics $M --clock-control none -k regex:remap_u8 -s 2 -c 1 python profiles/drive.py remap --side 1344 --batch 64 2>&1 | grep -E "gpu__time|smsp__|l1tex"
