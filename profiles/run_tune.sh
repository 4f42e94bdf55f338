cd $GRAFT_REPO_ROOT
for tr in 4 2 12 1; do
touch attwarp_b200/csrc/remap_stream.cu
ATTWARP_NVCC_EXTRA="-DAW_TILE_ROWS=$tr" python -m attwarp_b200.build > /dev/null 2>&1 || echo build failed
echo "== tile rows $tr"
timeout 120 python profiles/drive.py remap --side 336 --batch 256 --iters 12 | sed 's/GB.*//' | sed 's/.*us//'
timeout 120 python profiles/drive.py remap --side 1344 --batch 64 --iters 8| sed 's/GB.*//' | sed 's/.*us//'
done
touch attwarp_b200/csrc/remap_stream.cu; python -m attwarp_b200.build > /dev/null 2>&1
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
