# round 2, pass r09c: finish kernels with eight partial loads in flight
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
P=gpurun_out/r09c
timeout 1800 python -m pytest tests/ -x -q -m gpu -n 4 > ${P}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> ${P}_pytest_gpu.log; tail -n 3 ${P}_pytest_gpu.log | cut -c1-300
timeout 600 python profiles/row_kernels_probe.py 2>&1 | grep -v Warning > ${P}_row_kernels.txt; cat ${P}_row_kernels.txt
timeout 600 python profiles/formats_probe.py 2>&1 | grep -v Warning | grep "maps_from_attention" | grep "identity" > ${P}_formats_probe.txt; cat ${P}_formats_probe.txt
