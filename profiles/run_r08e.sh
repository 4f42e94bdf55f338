# round 2, pass r08e: tensor-core LANCZOS with the coefficient byte planes precomputed on the host
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
P=gpurun_out/r08e
timeout 900 python -m pytest tests/test_mask_path.py tests/test_gpu_stage_fuzz.py tests/test_save_warped_image.py tests/test_gpu_pipeline_fuzz.py -m gpu -q -n 4 > ${P}_pytest_mask.log 2>&1; echo "pytest exit $?" >> ${P}_pytest_mask.log; tail -n 12 ${P}_pytest_mask.log | cut -c1-600
timeout 600 python profiles/row_kernels_probe.py 2>&1 | grep -v Warning | grep -i "mota\|mask" > ${P}_row_kernels.txt; cat ${P}_row_kernels.txt
ATTWARP_LANCZOS_MMA=0 timeout 600 python profiles/row_kernels_probe.py 2>&1 | grep -v Warning | grep -i "mota_mask " > ${P}_row_kernels_imad.txt; cat ${P}_row_kernels_imad.txt
