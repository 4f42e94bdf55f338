# round 2, pass r06f: uint8 maps + transform through the row-owning kernel (256-entry table); walk kernel dropped; full GPU suite
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
P=gpurun_out/r06f
timeout 1500 python -m pytest tests -m gpu -q -n 4 > ${P}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> ${P}_pytest_gpu.log; tail -n 4 ${P}_pytest_gpu.log | cut -c1-400
timeout 600 python profiles/formats_probe.py 2>&1 | grep -v Warning | grep maps_from > ${P}_formats_probe.txt; cat ${P}_formats_probe.txt
timeout 600 python profiles/row_kernels_probe.py 2>&1 | grep -v Warning > ${P}_row_kernels.txt; cat ${P}_row_kernels.txt
