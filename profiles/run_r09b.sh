# round 2, pass r09b: ncu of the finish chain (maps_from_partials_kernel) at 64 x 1344^2 and 256 x 336^2
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
P=gpurun_out/r09b
timeout 600 ncu --set full --clock-control none --import-source on -k regex:maps_from_partials -s 2 -c 1 -o ${P}_prof_finish_1344 -f python profiles/drive.py att --side 1344 --batch 64 > ${P}_ncu.log 2>&1; tail -n 1 ${P}_ncu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:maps_from_partials -s 2 -c 1 -o ${P}_prof_finish_336 -f python profiles/drive.py att --side 336 --batch 256 > ${P}_ncu2.log 2>&1; tail -n 1 ${P}_ncu2.log
