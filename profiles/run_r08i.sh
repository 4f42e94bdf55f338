cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
P=gpurun_out/r08i
for m in 0 1 2 3; do echo "ATTWARP_HOST_PINNED=$m"; ATTWARP_HOST_PINNED=$m timeout 300 python profiles/single_image_probe.py 2>&1 | grep -v Warning | cut -c1-60; done > ${P}_pinned_modes.txt; cat ${P}_pinned_modes.txt
