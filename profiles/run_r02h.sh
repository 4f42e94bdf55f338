# round 2: raw CTA timeline at c2 (who finishes late?), c4 launch breakdown
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/r02h_trace_c2.txt
ATTWARP_REMAP_TRACE=gpurun_out/r02h_trace_c2.txt timeout 300 python profiles/s5_probe.py --only "c2  256x336^2 hwc near" --reps 2 > /dev/null 2>&1
tail -593 gpurun_out/r02h_trace_c2.txt > gpurun_out/r02h_trace_c2_last.txt; rm -f gpurun_out/r02h_trace_c2.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'maps_from|remap_u8' -c 24 --csv --log-file gpurun_out/r02h_c4_launches.csv python profiles/c4_probe.py --steps 2 > gpurun_out/r02h_c4_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02h_c4_launches.csv')) if len(r)>5]
hdr=rows[0]
kn=hdr.index('Kernel Name'); mv=hdr.index('Metric Value'); g=hdr.index('Grid Size'); b=hdr.index('Block Size')
for r in rows[-12:]:
    print(r[kn][:70], r[g], r[b], r[mv])
PY
timeout 300 python -m pytest tests/test_image_io.py -m gpu -q 2>&1 | tail -3
