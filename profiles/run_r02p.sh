# round 2: DIRECT-store mode of the uint8 quad kernel -- parity (all mapping policies), memcheck, timing A/B
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for pol in 0 1 2; do
ATTWARP_QUAD_MAP=$pol timeout 900 python -m pytest tests/test_gpu_remap_edges.py tests/test_gpu_ragged.py tests/test_gpu_fused_batch.py tests/test_gpu_numpy_path.py tests/test_save_warped_image.py -m gpu -q -x > gpurun_out/r02p_pytest_map$pol.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02p_pytest_map$pol.log
tail -5 gpurun_out/r02p_pytest_map$pol.log
done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_remap_edges.py tests/test_gpu_ragged.py -m gpu -q -x -k "not minification" > gpurun_out/r02p_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r02p_memcheck.log
tail -4 gpurun_out/r02p_memcheck.log
echo "== direct"; timeout 300 python profiles/s5_probe.py --only c --reps 40 2>&1 | tee gpurun_out/r02p_probe_direct.txt
echo "== tiles"; ATTWARP_QUAD_DIRECT=0 timeout 300 python profiles/s5_probe.py --only c --reps 40 2>&1 | tee gpurun_out/r02p_probe_tiles.txt
echo "== direct all"; timeout 300 python profiles/s5_probe.py --reps 40 2>&1 | tee gpurun_out/r02p_probe_direct_all.txt
for r in 8 12 16; do echo "== direct rows $r"; ATTWARP_QUAD_ROWS=$r timeout 300 python profiles/s5_probe.py --only c --reps 40 2>&1 | tee gpurun_out/r02p_probe_direct_r$r.txt; done
echo "== direct ring 3"; ATTWARP_QUAD_RING=32 timeout 300 python profiles/s5_probe.py --only c --reps 40 2>&1 | tee gpurun_out/r02p_probe_direct_ring3.txt
echo "== c4 direct"; timeout 300 python profiles/c4_probe.py 2>&1 | tee gpurun_out/r02p_c4_direct.txt
echo "== c4 direct round 4"; timeout 300 python profiles/c4_probe.py --round 4 2>&1 | tee gpurun_out/r02p_c4_direct_r4.txt
echo "== c4 tiles"; ATTWARP_QUAD_DIRECT=0 timeout 300 python profiles/c4_probe.py 2>&1 | tee gpurun_out/r02p_c4_tiles.txt
echo "== wait hint"; ATTWARP_QUAD_WAIT_HINT=2000 timeout 300 python profiles/s5_probe.py --only c --reps 40 2>&1 | tee gpurun_out/r02p_probe_hint.txt
echo "== roles first"; ATTWARP_QUAD_ROLES_FIRST=1 timeout 300 python profiles/s5_probe.py --only c --reps 40 2>&1 | tee gpurun_out/r02p_probe_roles.txt
