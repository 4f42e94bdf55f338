# round 2: class-merge threshold for small ragged batches (an 8-GPU shard is 128 images); secondary kernels by graph replay
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for px in 2000000 8000000 16000000 32000000; do
echo "== ATTWARP_QUAD_CLASS_MIN_PX=$px"
ATTWARP_QUAD_CLASS_MIN_PX=$px timeout 300 python profiles/c4_probe.py --n 128 2>&1 | tail -1
ATTWARP_QUAD_CLASS_MIN_PX=$px timeout 300 python profiles/c4_probe.py --n 256 2>&1 | tail -1
ATTWARP_QUAD_CLASS_MIN_PX=$px timeout 300 python profiles/c4_probe.py --n 1024 2>&1 | tail -1
done 2>&1 | tee gpurun_out/r03c_c4_class_min_px.txt
timeout 600 python profiles/row_kernels_probe.py 2>&1 | tee gpurun_out/r03c_row_kernels.txt
