cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:remap_u8 -s 5 -c 1 -o gpurun_out/r03h_prof_c3ud -f python profiles/s5_probe.py --only c3ud --reps 4 > gpurun_out/r03h_ncu_c3ud.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:remap_u8 -s 5 -c 1 -o gpurun_out/r03h_prof_c3 -f python profiles/s5_probe.py --only "c3 " --reps 4 > gpurun_out/r03h_ncu_c3.log 2>&1
tail -2 gpurun_out/r03h_ncu_c3ud.log gpurun_out/r03h_ncu_c3.log
