# round 2: ncu --set full with source counters on the quad kernel (c3 and c2 shapes)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:remap_u8_quad -s 2 -c 1 -o gpurun_out/r02d_quad1344 -f python profiles/drive.py remap --side 1344 --batch 64 --grid 48 > gpurun_out/r02d_ncu1344.log 2>&1
tail -3 gpurun_out/r02d_ncu1344.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:remap_u8_quad -s 2 -c 1 -o gpurun_out/r02d_quad336 -f python profiles/drive.py remap --side 336 --batch 256 > gpurun_out/r02d_ncu336.log 2>&1
tail -3 gpurun_out/r02d_ncu336.log
ls -la gpurun_out/*.ncu-rep
