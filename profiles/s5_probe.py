#!/usr/bin/env python
"""Stage-5 (uint8 resample) timing probe: the kernel alone at the bench shapes, back-to-back launches over rotating
buffers, optionally with the kernel's debug switch (ATTWARP_REMAP_DBG=1: skip the sweep -> what the load pipeline +
launch ramp cost without the arithmetic and the stores).  Never a bench number.

    python profiles/s5_probe.py [--dbg]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from attwarp_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dbg", action="store_true")
ap.add_argument("--reps", type=int, default=30)
ap.add_argument("--only", default="")
args = ap.parse_args()
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(0)
PEAK = 6560.3


def maps(B, side, grid, kind):
    if kind == "c2":          # near-uniform token maps (mean of 1024 softmax rows)
        tok = 1.0 + 0.03 * torch.randn(B, grid, grid, device=dev, generator=gen)
    else:
        tok = torch.rand(B, grid, grid, device=dev, generator=gen) ** 3
    tok = (tok / tok.sum(dim=(1, 2), keepdim=True)).contiguous()
    return ops.maps_from_tokens(tok, (side, side))


def run(name, B, side, grid, kind, layout="hwc", C=3, R=4, wside=None, woside=None):
    if args.only and not name.strip().startswith(args.only):
        return
    wside = side if wside is None else wside          # source width
    woside = wside if woside is None else woside      # output width
    shape = (B, side, wside, C) if layout == "hwc" else (B, C, side, wside)
    oshape = (B, side, woside, C) if layout == "hwc" else (B, C, side, woside)
    imgs = [torch.randint(0, 256, shape, device=dev, dtype=torch.uint8, generator=gen) for _ in range(R)]
    outs = [torch.empty(oshape, device=dev, dtype=torch.uint8) for _ in imgs]
    if wside != side or woside != side:
        tok = torch.rand(B, grid, grid, device=dev, generator=gen) ** 3
        mx, my = ops.maps_from_tokens((tok / tok.sum(dim=(1, 2), keepdim=True)).contiguous(), (side, wside), (side, woside))
    else:
        mx, my = maps(B, side, grid, kind)
    by = imgs[0].numel() + outs[0].numel()
    for dbg in (["0", "1"] if args.dbg else ["0"]):
        os.environ["ATTWARP_REMAP_DBG"] = dbg
        for i in range(3):
            ops.remap_bilinear(imgs[i % R], mx, my, layout, out=outs[i % R])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.reps):
            ops.remap_bilinear(imgs[i % R], mx, my, layout, out=outs[i % R])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        print(f"{name:34s} dbg={dbg}  {ms * 1e3:8.1f} us  {by / ms / 1e6:7.0f} GB/s  frac {by / ms / 1e6 / PEAK:.3f}", flush=True)
    os.environ["ATTWARP_REMAP_DBG"] = "0"
    del imgs, outs
    torch.cuda.empty_cache()


run("c2  256x336^2 hwc near-uniform", 256, 336, 24, "c2", R=8)
run("c2  256x336^2 hwc rand^3", 256, 336, 24, "c3", R=8)
run("c2u 256x336x335 rows unaligned", 256, 336, 24, "c3", R=8, wside=335)
run("c3u 64x1344x1343 rows unaligned", 64, 1344, 48, "c3", wside=1343)
run("c3us 64x1344: 1343 -> 1344 (source rows unaligned only)", 64, 1344, 48, "c3", wside=1343, woside=1344)
run("c3ud 64x1344: 1344 -> 1343 (destination rows unaligned only)", 64, 1344, 48, "c3", wside=1344, woside=1343)
run("    1024x336^2 hwc near-uniform", 1024, 336, 24, "c2", R=3)
run("c3  64x1344^2 hwc rand^3 48x48", 64, 1344, 48, "c3")
run("    256x1344^2 hwc rand^3 48x48", 256, 1344, 48, "c3", R=2)
run("    256x3x336^2 chw rand^3", 256, 336, 24, "c3", layout="chw", R=8)
run("    256x336^2x4 hwc rand^3", 256, 336, 24, "c3", C=4, R=8)
run("    256x336^2x1 hwc rand^3", 256, 336, 24, "c3", C=1, R=8)
