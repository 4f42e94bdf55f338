# round 2: ncu --set full with source counters on the quad kernel (c3 and c2 shapes); c4 with width classes
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ragged.py -m gpu -q -x > gpurun_out/r02d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02d_pytest.log; tail -3 gpurun_out/r02d_pytest.log
timeout 300 python profiles/c4_probe.py > gpurun_out/r02d_c4.txt 2>&1
ATTWARP_QUAD_GEO=2 timeout 300 python profiles/c4_probe.py >> gpurun_out/r02d_c4.txt 2>&1
ATTWARP_REMAP_QUAD=0 timeout 300 python profiles/c4_probe.py >> gpurun_out/r02d_c4.txt 2>&1
cat gpurun_out/r02d_c4.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:remap_u8_quad -s 2 -c 1 -o gpurun_out/r02d_quad1344 -f python profiles/drive.py remap --side 1344 --batch 64 --grid 48 > gpurun_out/r02d_ncu1344.log 2>&1
tail -3 gpurun_out/r02d_ncu1344.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:remap_u8_quad -s 2 -c 1 -o gpurun_out/r02d_quad336 -f python profiles/drive.py remap --side 336 --batch 256 > gpurun_out/r02d_ncu336.log 2>&1
tail -3 gpurun_out/r02d_ncu336.log
ls -la gpurun_out/*.ncu-rep
