# round 2, pass r09a: ncu of the float32 marginals kernel (identity, 64 x 1344^2) -- what bounds it
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
P=gpurun_out/r09a
timeout 600 ncu --set full --clock-control none --import-source on -k regex:marginals_f32_rows -s 2 -c 1 -o ${P}_prof_marginals_f32 -f python profiles/drive.py att --side 1344 --batch 64 --dtype f32 > ${P}_ncu.log 2>&1; tail -n 2 ${P}_ncu.log
