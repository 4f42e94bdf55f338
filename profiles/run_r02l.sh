# round 2: what costs c4 -- alignment, two-strip images, or the mix?
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for args in "--round 1" "--round 16" "--round 4" "--round 1 --min-side 705 --max-side 1408" "--round 16 --min-side 705 --max-side 1408" "--round 1 --min-side 1409 --max-side 2048" "--round 16 --min-side 1409 --max-side 2048" "--round 16 --min-side 1344 --max-side 1344 --n 256"; do
timeout 300 python profiles/c4_probe.py $args >> gpurun_out/r02l_c4.txt 2>&1
done
cat gpurun_out/r02l_c4.txt
