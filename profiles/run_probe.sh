cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python profiles/drive.py att --side 336 --batch 256 --iters 6
python profiles/drive.py att --side 1344 --batch 64 --iters 6
python profiles/drive.py att --side 512 --batch 128 --dtype f32 --iters 6
M=gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum
for cfg in "--side 1344 --batch 64" "--side 336 --batch 256" "--side 512 --batch 128 --dtype f32"; do
timeout 300 ncu --metrics $M --clock-control none -k regex:'marginals' -s 2 -c 1 python profiles/drive.py att $cfg --iters 3 2>&1 | grep -E "gpu__time|smsp__|dram" | awk '{printf "%s ", $NF} END {print ""}'
done
