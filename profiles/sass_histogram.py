#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libattwarp_sm100.so (cuobjdump -sass): the evidence that the hot path is
hand-written sm_100a code -- UBLKCP (cp.async.bulk, the TMA bulk copy), SYNCS (mbarrier), IDP (dp4a / dp2a),
IMAD, PRMT, SHF, FFMA2 / FADD2 (packed fp32), DFMA / DADD (float64 stages 2-4), LDS / STS widths.

    python profiles/sass_histogram.py [path/to/lib.so] > profiles/r02_sass_histogram.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "attwarp_b200", "libattwarp_sm100.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n  # noqa: E731

kernels, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", line)
    if m and cur:
        kernels[cur][m.group(1)] += 1

KEYS = ["UBLKCP", "SYNCS", "IMMA", "IDP.4A", "IDP.2A", "IMAD", "PRMT", "SHF", "LOP3", "FFMA2", "FADD2", "FMUL2", "DFMA", "DADD",
        "DMUL", "LDS", "STS", "STS.U8", "LDG", "STG", "SHFL", "REDUX", "BAR", "BRA"]


def family(counter, key):
    if key in ("STS.U8", "IDP.4A", "IDP.2A"):
        return sum(v for k, v in counter.items() if k.startswith(key))
    return sum(v for k, v in counter.items() if k == key or k.startswith(key + "."))


print("# SASS opcode histogram per kernel (`cuobjdump -sass attwarp_b200/libattwarp_sm100.so`)\n")
print("Static instruction counts (not executed counts).  `UBLKCP` = cp.async.bulk (TMA bulk copy), `SYNCS` = mbarrier, "
      "`IDP` = dp4a/dp2a, `IMMA` = integer mma.sync (the LANCZOS vertical pass).\n")
print("| kernel | total | " + " | ".join(KEYS) + " |")
print("|---|---:|" + "---:|" * len(KEYS))
for name, c in kernels.items():
    short = demangle(name)
    short = re.sub(r"\(anonymous namespace\)::|aw::|void ", "", short)
    short = re.sub(r"\((int|bool|unsigned int)\)", "", short)
    short = re.sub(r"\(.*$", "", short)
    print(f"| `{short}` | {sum(c.values())} | " + " | ".join(str(family(c, k)) for k in KEYS) + " |")
