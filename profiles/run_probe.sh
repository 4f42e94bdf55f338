cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
for cpt in 2 1; do
echo "CPT $cpt: chw C=3 336 x256; hwc C=1 1344 x64; chw C=3 333 x256"
ATTWARP_REMAP_CPT=$cpt timeout 300 ncu --metrics $M --clock-control none -k regex:remap_u8 -s 2 -c 1 python profiles/drive.py remap --side 336 --batch 256 --layout chw --iters 3 2>&1 | grep -E "gpu__time|smsp__" | awk '{printf "%s ", $NF} END {print ""}'
ATTWARP_REMAP_CPT=$cpt timeout 300 ncu --metrics $M --clock-control none -k regex:remap_u8 -s 2 -c 1 python profiles/drive.py remap --side 1344 --batch 64 --C 1 --iters 3 2>&1 | grep -E "gpu__time|smsp__" | awk '{printf "%s ", $NF} END {print ""}'
ATTWARP_REMAP_CPT=$cpt timeout 300 ncu --metrics $M --clock-control none -k regex:remap_u8 -s 2 -c 1 python profiles/drive.py remap --side 333 --batch 256 --layout chw --iters 3 2>&1 | grep -E "gpu__time|smsp__" | awk '{printf "%s ", $NF} END {print ""}'
done
