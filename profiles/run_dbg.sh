cd $GRAFT_REPO_ROOT
for d in 0 1 2 3; do
echo "== dbg $d"
ATTWARP_REMAP_DBG=$d timeout 120 python profiles/drive.py remap --side 336 --batch 256 --iters 8 | sed 's/GB.*//'
ATTWARP_REMAP_DBG=$d timeout 120 python profiles/drive.py remap --side 1344 --batch 64 --iters 6 | sed 's/GB.*//'
done
