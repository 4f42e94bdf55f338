# round 2, pass r08h: the single-image host entry point stages the caller's buffers through pinned arena memory
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
P=gpurun_out/r08h
timeout 900 python -m pytest tests/test_gpu_numpy_path.py tests/test_save_warped_image.py tests/test_mask_path.py tests/test_gpu_pipeline_fuzz.py -m gpu -q -n 4 > ${P}_pytest.log 2>&1; echo "pytest exit $?" >> ${P}_pytest.log; tail -n 4 ${P}_pytest.log | cut -c1-400
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
(timeout 300 python profiles/single_image_probe.py; echo "ATTWARP_HOST_PINNED=0:"; ATTWARP_HOST_PINNED=0 timeout 300 python profiles/single_image_probe.py; timeout 300 python profiles/pinned_staging_probe.py) 2>&1 | grep -v Warning > ${P}_single_image_probe.txt; cat ${P}_single_image_probe.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_numpy_path.py -m gpu -q -x -k "drop_in or warp_image or host" > ${P}_memcheck.log 2>&1; echo "memcheck exit $?" >> ${P}_memcheck.log; tail -n 3 ${P}_memcheck.log | cut -c1-300
