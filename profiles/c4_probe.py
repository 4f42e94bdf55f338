#!/usr/bin/env python
"""configs[3] (1024 mixed-resolution images) on one GPU: host time of the Python call vs GPU time of a step.
Never a bench number.    python profiles/c4_probe.py [--n 1024]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from attwarp_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1024)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--round", type=int, default=1, help="round the sides down to a multiple of this")
ap.add_argument("--odd", action="store_true", help="make every side odd (all rows unaligned)")
ap.add_argument("--max-side", type=int, default=2048)
ap.add_argument("--min-side", type=int, default=224)
a = ap.parse_args()
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
sides = np.random.default_rng(1237).integers(a.min_side, a.max_side + 1, size=a.n)
sides = (sides // a.round) * a.round
if a.odd:
    sides = sides | 1        # every width odd: no destination row of any image is 4-byte aligned
imgs = [torch.randint(0, 256, (int(s), int(s), 3), device=dev, dtype=torch.uint8, generator=g) for s in sides]
outs = [torch.empty_like(i) for i in imgs]
tok = torch.rand(a.n, 24, 24, device=dev, generator=g) ** 3
tok = (tok / tok.sum(dim=(1, 2), keepdim=True)).contiguous()
by = 2 * sum(t.numel() for t in imgs)
for _ in range(3):
    ops.warp_ragged_from_tokens(tok, imgs, outs=outs)
torch.cuda.synchronize()
# host time per call (GPU idle at the start of each call)
hs = []
for _ in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ops.warp_ragged_from_tokens(tok, imgs, outs=outs)
    hs.append(time.perf_counter() - t0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    ops.warp_ragged_from_tokens(tok, imgs, outs=outs)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
print(f"[round {a.round}, sides {a.min_side}..{a.max_side}] ", end="")
print(f"c4 n={a.n}: step {ms:.3f} ms  {by / ms / 1e6:.0f} GB/s  frac {by / ms / 1e6 / 6560.3:.3f}   host call {min(hs) * 1e3:.3f} ms "
      f"(QUAD={os.environ.get('ATTWARP_REMAP_QUAD', '1')})", flush=True)
if hasattr(ops, "RaggedBatch"):
    rb = ops.RaggedBatch(imgs, outs=outs)
    for _ in range(3):
        rb.run(tok)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rb.run(tok)
    h = time.perf_counter() - t0
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.steps):
        rb.run(tok)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(f"c4 n={a.n} RaggedBatch: step {ms:.3f} ms  {by / ms / 1e6:.0f} GB/s  frac {by / ms / 1e6 / 6560.3:.3f}   host call {h * 1e3:.3f} ms", flush=True)
