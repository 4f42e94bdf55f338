# round 2, pass r08d: hooked attention step through a prepared C call (host-side cost per step)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
P=gpurun_out/r08d
timeout 900 python -m pytest tests/test_gpu_torch_path.py -m gpu -q -k "hook" > ${P}_pytest.log 2>&1; echo "pytest exit $?" >> ${P}_pytest.log; tail -n 6 ${P}_pytest.log | cut -c1-500
(timeout 300 python profiles/hook_overhead.py; ATTWARP_HOOK_FAST=0 timeout 300 python profiles/hook_overhead.py) 2>&1 | grep -v Warning > ${P}_hook_overhead.txt; cat ${P}_hook_overhead.txt
