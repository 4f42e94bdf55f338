# round 2, pass r06d: ncu of the walk kernel (single-channel 256 x 336^2) -- why 53 us
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
P=gpurun_out/r06d
timeout 600 ncu --set full --clock-control none --import-source on -k regex:remap_u8_walk -s 2 -c 1 -o ${P}_prof_walk_c1 -f python profiles/drive.py remap --side 336 --batch 256 --C 1 > ${P}_ncu_walk_c1.log 2>&1; tail -n 3 ${P}_ncu_walk_c1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:remap_u8_walk -s 2 -c 1 -o ${P}_prof_walk_c4 -f python profiles/drive.py remap --side 336 --batch 256 --C 4 > ${P}_ncu_walk_c4.log 2>&1; tail -n 3 ${P}_ncu_walk_c4.log
