#!/usr/bin/env python
"""Digest of one kernel of an .ncu-rep: headline metrics, stall mix, and the SASS hot spots.

    python profiles/ncu_digest.py gpurun_out/prof.ncu-rep [kernel-index]
"""
import csv, io, re, subprocess, sys

def page(rep, name):
    return list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"],
            stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout)))

def main():
    rep = sys.argv[1]
    idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = page(rep, "raw")
    hdr, units, r = rows[0], rows[1], rows[2 + idx]
    want = re.compile(r"^(gpu__time_duration.sum|dram__bytes_(read|write).sum|smsp__inst_executed.sum|"
                      r"smsp__issue_active.avg.pct|sm__warps_active.avg.pct|"
                      r"smsp__average_warps_issue_stalled_.*_per_issue_active|"
                      r"l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum|"
                      r"sm__inst_executed_pipe_(alu|fma|fmaheavy|lsu|xu|uniform).*sum.pct_of_peak_sustained_active|"
                      r"launch__(registers_per_thread$|grid_size|block_size|shared_mem_per_block_dynamic)|"
                      r"sm__cycles_elapsed.max$|l1tex__data_pipe_lsu_wavefronts.avg.pct|gpu__dram_throughput.avg.pct|"
                      r"lts__t_sectors_op_(read|write).sum$|lts__throughput.avg.pct)")
    print(r[hdr.index("Kernel Name")][:100])
    for h, u, v in zip(hdr, units, r):
        if want.search(h) and not (("stalled" in h) and float(v.replace(",", "")) < 0.05):
            print(f"  {h:88s} {v} {u}")
    rows = page(rep, "source")
    hdr = rows[1]
    iS, iSt, iI = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    data = rows[2:]
    tot = sum(int(r[iI]) for r in data); tots = sum(int(r[iSt]) for r in data)
    print("SASS: total warp instructions", tot, "stall samples", tots)
    for k, r in enumerate(data):
        if int(r[iSt]) > tots * 0.015 or int(r[iI]) > tot * 0.03:
            print(f"  {k:5d} n={int(r[iI]):9d} ({100*int(r[iI])/tot:4.1f}%) stall={100*int(r[iSt])/tots:5.2f}%  {r[iS][:80]}")

if __name__ == "__main__":
    main()
