cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python profiles/kernel_survey.py 2>&1 | tee gpurun_out/kernel_survey.txt | grep -E "pool|gt_marg"
