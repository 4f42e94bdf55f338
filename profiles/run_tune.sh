cd $GRAFT_REPO_ROOT
for cfg in "3 2 12 3" "3 2 8 4" "2 2 12 4" "3 2 6 5"; do
set -- $cfg
touch attwarp_b200/csrc/remap_stream.cu
ATTWARP_NVCC_EXTRA="-DAW_SRC_STAGES=$1 -DAW_OUT_STAGES=$2 -DAW_ROWS=$3 -DAW_MIN_CTAS=$4" python -m attwarp_b200.build > /dev/null 2>&1 || echo build failed
echo "== S=$1 O=$2 R=$3 CTAS=$4"
timeout 120 python profiles/drive.py remap --side 336 --batch 256 --iters 12 | sed 's/GB.*//' | sed 's/.*us//'
timeout 120 python profiles/drive.py remap --side 1344 --batch 64 --iters 8| sed 's/GB.*//' | sed 's/.*us//'
done
