cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
for c in 2 1; do
echo "== cpt $c"
ATTWARP_REMAP_CPT=$c timeout 120 python profiles/drive.py remap --side 336 --batch 256 --iters 10 | sed 's/GB.*//'
ATTWARP_REMAP_CPT=$c timeout 120 python profiles/drive.py remap --side 1344 --batch 64 --iters 10| sed 's/GB.*//'
ATTWARP_REMAP_CPT=$c timeout 120 python profiles/drive.py remap --side 336 --out-side 500 --batch 256 --iters 5| sed 's/GB.*//'
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:remap_u8 -s 2 -c 1 -o gpurun_out/prof_stream336 -f python profiles/drive.py remap --side 336 --batch 256 > gpurun_out/ncu_stream336.log 2>&1
