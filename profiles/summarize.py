#!/usr/bin/env python
"""Turn ncu captures brought back in gpurun_out/ into the small tracked summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_c2.csv  profiles/r01_c2_launches.md
    python profiles/summarize.py full     gpurun_out/prof_c2.ncu-rep  profiles/r01_c2_ncu_full.md  [c2]

`launches`: the `--metrics gpu__time_duration.sum` pass -> per-kernel count / mean / share of the
step among this library's kernels (namespace aw::).  Times under ncu are cold-cache and serialised:
the SHARES are what must agree with bench.py, not the absolute values.
`full`: one `--set full` capture -> per-kernel DRAM bytes, throughput percentages, occupancy,
registers; with a workload key it also records per-launch DRAM traffic in profiles/traffic.json
(read by bench.py for `roofline.traffic`).
"""

import csv
import io
import json
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def short(name):
    m = re.search(r"(\w+_kernel)(<[^>]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name[:60]


def launches(path, out):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, ib, ig = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Block Size", "Grid Size"))
    ours, others = {}, 0.0
    seq = []
    for r in rows[1:]:
        ns = float(r[iv].replace(",", ""))
        if "aw::" in r[ik] or "unnamed>::" in r[ik]:
            k = short(r[ik])
            d = ours.setdefault(k, {"n": 0, "ns": 0.0, "grid": r[ig], "block": r[ib]})
            d["n"] += 1
            d["ns"] += ns
            seq.append((k, ns))
        else:
            others += ns
    tot = sum(d["ns"] for d in ours.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list: {os.path.basename(path)}\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised "
                "launches: compare shares, not absolutes).\n\n")
        f.write("| kernel (aw::) | launches | mean us | share of library time | grid | block |\n|---|---:|---:|---:|---|---|\n")
        for k, d in sorted(ours.items(), key=lambda kv: -kv[1]["ns"]):
            f.write(f"| `{k}` | {d['n']} | {d['ns'] / d['n'] / 1e3:.2f} | {100 * d['ns'] / tot:.1f} % | {d['grid']} | {d['block']} |\n")
        f.write(f"\nOther (torch input generation etc.): {others / 1e3:.1f} us total, not part of the timed step.\n")
        f.write("\nLast launches in order (us): " + ", ".join(f"{k.split('_kernel')[0]} {ns / 1e3:.1f}" for k, ns in seq[-9:]) + "\n")
    print(open(out).read())


WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2 % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
]


def to_bytes(val, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(val.replace(",", "")) * mult.get(unit, 1)


def full(path, out, workload=None):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    traffic = {}
    with open(out, "w") as f:
        f.write(f"# ncu --set full: {os.path.basename(path)}\n\n")
        for r in rows[2:]:
            k = short(r[ik])
            f.write(f"## `{k}`\n\n| metric | value |\n|---|---|\n")
            for m, label in WANT:
                if m in hdr:
                    i = hdr.index(m)
                    f.write(f"| {label} (`{m}`) | {r[i]} {units[i]} |\n")
            try:
                ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
                tb = to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw])
                traffic[k.split("<")[0]] = tb
                f.write(f"| **DRAM traffic per launch** | {tb / 1e6:.2f} MB |\n")
            except ValueError:
                pass
            f.write("\n")
    if workload:
        tp = os.path.join(HERE, "traffic.json")
        try:
            cur = json.load(open(tp))
        except Exception:
            cur = {}
        cur.setdefault(workload, {}).update(traffic)
        json.dump(cur, open(tp, "w"), indent=1, sort_keys=True)
    print(open(out).read())


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
