cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { echo "== $1"; env $1 timeout 300 python profiles/s5_probe.py --only "c" --reps 40 2>&1 | grep -v "c2u\|c3u"; }
run "X=0"
run "ATTWARP_QUAD_FIRST_ROWS=4"
run "ATTWARP_QUAD_FIRST_ROWS=6"
run "ATTWARP_QUAD_FIRST_ROWS=8"
run "ATTWARP_QUAD_FIRST_ROWS=12"
