# experiment: QUAD mapping, aligned rows: each lane stores its own 12 bytes (three 32-bit stores, 12-byte lane stride)
# instead of exchanging them through shared memory for 128-byte-contiguous warp stores
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_remap_edges.py -m gpu -q -x -k "mappings or alignment" > gpurun_out/r02w_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02w_pytest.log
tail -4 gpurun_out/r02w_pytest.log
for p in 1 0; do echo "== policy $p"; ATTWARP_QUAD_MAP=$p timeout 300 python profiles/s5_probe.py --only c --reps 40 2>&1 | tee gpurun_out/r02w_probe_map$p.txt; done
ATTWARP_QUAD_MAP=1 timeout 300 python profiles/c4_probe.py --round 4 2>&1 | tee gpurun_out/r02w_c4_r4.txt
