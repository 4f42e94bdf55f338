cd $GRAFT_REPO_ROOT
for cfg in "32 1" "16 2" "16 1" "24 2" "48 1"; do
set -- $cfg
touch attwarp_b200/csrc/remap.cu
ATTWARP_NVCC_EXTRA="-DAW_F32_ROWS=$1 -DAW_F32_COLS=$2" python -m attwarp_b200.build > /dev/null 2>&1 || echo build failed
echo "== rows=$1 cols=$2"
timeout 120 python profiles/drive.py remap --side 512 --batch 128 --dtype f32 --layout chw --iters 8 | sed 's/.*GB.s//'
timeout 120 python profiles/drive.py remap --side 512 --batch 128 --dtype f32 --layout hwc --iters 8 | sed 's/.*GB.s//'
done
