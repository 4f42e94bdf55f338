cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:remap_u8 -s 2 -c 1 -o gpurun_out/prof_stream336 -f python profiles/drive.py remap --side 336 --batch 256 > gpurun_out/ncu_stream336.log 2>&1
tail -3 gpurun_out/ncu_stream336.log
