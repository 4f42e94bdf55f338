# round 2: CTA timeline of the quad kernel at c2, c4 launch breakdown, new tests (image_io, fused softmax+mix, mappings)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_remap_edges.py tests/test_image_io.py tests/test_gpu_autograd.py tests/test_pdf_loss.py -m gpu -q > gpurun_out/r02g_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02g_pytest.log
tail -25 gpurun_out/r02g_pytest.log
rm -f gpurun_out/r02g_trace_c2.txt
ATTWARP_REMAP_TRACE=gpurun_out/r02g_trace_c2.txt timeout 300 python profiles/s5_probe.py --only "c2  256x336^2 hwc near" --reps 4 > gpurun_out/r02g_trace_probe.txt 2>&1
python profiles/trace_digest.py gpurun_out/r02g_trace_c2.txt > gpurun_out/r02g_trace_c2_digest.txt 2>&1; cat gpurun_out/r02g_trace_c2_digest.txt
rm -f gpurun_out/r02g_trace_c3.txt
ATTWARP_REMAP_TRACE=gpurun_out/r02g_trace_c3.txt timeout 300 python profiles/s5_probe.py --only "c3" --reps 4 > /dev/null 2>&1
python profiles/trace_digest.py gpurun_out/r02g_trace_c3.txt > gpurun_out/r02g_trace_c3_digest.txt 2>&1; cat gpurun_out/r02g_trace_c3_digest.txt
rm -f gpurun_out/r02g_trace_c2.txt gpurun_out/r02g_trace_c3.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02g_c4_launches.csv python profiles/c4_probe.py --steps 2 > gpurun_out/r02g_c4_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02g_c4_launches.csv')) if len(r)>5]
hdr=rows[0]
kn=hdr.index('Kernel Name'); mv=hdr.index('Metric Value')
for r in rows[-12:]:
    print(r[kn][:90], r[mv])
PY
