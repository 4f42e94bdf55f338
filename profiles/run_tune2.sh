# A/B of stage-5 build options on the GPU box (ncu durations + instruction counts)
cd $GRAFT_REPO_ROOT
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
for opt in "-DAW_MIN_CTAS=4 -DAW_SRC_STAGES=2 -DAW_ROWS=12" "-DAW_MIN_CTAS=4 -DAW_SRC_STAGES=2 -DAW_ROWS=10" "-DAW_MIN_CTAS=4 -DAW_SRC_STAGES=3 -DAW_ROWS=8" "-DAW_SRC_STAGES=2 -DAW_ROWS=16 -DAW_TILE_ROWS=4"; do
touch attwarp_b200/csrc/remap_stream.cu
ATTWARP_NVCC_EXTRA="$opt" python -m attwarp_b200.build > /dev/null 2>&1 || echo build failed
echo "== [$opt]"
for cfg in "--side 336 --batch 256" "--side 1344 --batch 64"; do
timeout 300 ncu --metrics $M --clock-control none -k regex:remap_u8 -s 2 -c 3 python profiles/drive.py remap $cfg --iters 4 2>&1 | grep -E "gpu__time|smsp__inst_exec" | awk '{printf "%s ", $NF} END {print ""}'
done
done
touch attwarp_b200/csrc/remap_stream.cu; python -m attwarp_b200.build > /dev/null 2>&1
