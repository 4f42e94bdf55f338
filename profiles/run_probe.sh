cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
for opt in "" "-DAW_F32_COLS=4" "-DAW_F32_COLS=1" "-DAW_F32_ROWS=16"; do
touch attwarp_b200/csrc/remap.cu
ATTWARP_NVCC_EXTRA="$opt" python -m attwarp_b200.build > /dev/null 2>&1 || echo build failed
echo "== [$opt]"
for lay in chw hwc; do
timeout 300 ncu --metrics $M --clock-control none -k regex:remap_f32 -s 2 -c 2 python profiles/drive.py remap --side 512 --batch 128 --dtype f32 --layout $lay --iters 3 2>&1 | grep -E "gpu__time|smsp__" | awk '{printf "%s ", $NF} END {print ""}'
done
done
touch attwarp_b200/csrc/remap.cu; python -m attwarp_b200.build > /dev/null 2>&1
