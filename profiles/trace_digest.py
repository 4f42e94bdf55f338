#!/usr/bin/env python
"""Digest of ATTWARP_REMAP_TRACE files: per launch, when (relative to the earliest CTA start of the launch) the CTAs
started, issued their first copy, saw their first rows, finished their first chunk, shipped their first and last tile.
    python profiles/trace_digest.py trace.txt"""
import sys

import numpy as np

launches, cur = [], None
for line in open(sys.argv[1]):
    if line.startswith("launch"):
        cur = {"hdr": line.strip(), "rows": []}
        launches.append(cur)
    else:
        cur["rows"].append([int(x) for x in line.split()[1:]])
names = ["cta start", "first copy issued", "first rows landed", "first chunk swept", "first tile shipped",
         "last tile shipped", "last chunk swept"]
for L in launches[-3:]:
    a = np.array(L["rows"], dtype=np.float64)
    t0 = a[:, 0].min()
    print(L["hdr"])
    for k, nm in enumerate(names):
        v = (a[:, k] - t0) / 1e3
        v = v[a[:, k] > 0]
        if len(v):
            print(f"  {nm:20s} min {v.min():7.2f}  p50 {np.median(v):7.2f}  p90 {np.percentile(v, 90):7.2f}  max {v.max():7.2f} us")
