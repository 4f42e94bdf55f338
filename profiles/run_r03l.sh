cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ragged.py -m gpu -q -x > gpurun_out/r03l_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r03l_pytest.log; tail -3 gpurun_out/r03l_pytest.log
(for n in 128 256 1024; do timeout 300 python profiles/c4_probe.py --n $n 2>&1 | tail -1; done; timeout 300 python profiles/c4_probe.py --round 4 2>&1 | tail -1) | tee gpurun_out/r03l_c4.txt
