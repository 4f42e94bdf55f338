# round 2: several store warps for wide strips
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_remap_edges.py tests/test_gpu_ragged.py tests/test_gpu_fused_batch.py tests/test_gpu_numpy_path.py -m gpu -q -x > gpurun_out/r02k_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02k_pytest.log
tail -4 gpurun_out/r02k_pytest.log
timeout 300 python profiles/s5_probe.py --only c > gpurun_out/r02k_probe.txt 2>&1; cat gpurun_out/r02k_probe.txt
timeout 300 python profiles/c4_probe.py > gpurun_out/r02k_c4.txt 2>&1; cat gpurun_out/r02k_c4.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'maps_from|remap_u8' -c 8 --csv --log-file gpurun_out/r02k_c4_launches.csv python profiles/c4_probe.py --steps 1 > gpurun_out/r02k_c4_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02k_c4_launches.csv')) if len(r)>5]
hdr=rows[0]
kn=hdr.index('Kernel Name'); mv=hdr.index('Metric Value'); g=hdr.index('Grid Size'); b=hdr.index('Block Size')
for r in rows[-4:]:
    print(r[kn][:70], r[g], r[b], r[mv])
PY
