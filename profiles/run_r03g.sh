cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python profiles/s5_probe.py --only c3 --reps 40 2>&1 | tee gpurun_out/r03g_probe.txt
