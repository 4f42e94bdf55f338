#!/usr/bin/env python
"""Device time of stage 5 on the formats that are not the benchmark's (grey, 4-channel, planar uint8) and of the
image-resolution marginals per transform and dtype, each replayed back to back from a CUDA graph over rotating
inputs (L2 cannot hold a rotation).  Never a bench number.

    python profiles/formats_probe.py            # ATTWARP_U8_WALK=0 for the round-1 kernel
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from attwarp_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(0)
PEAK = 6560.3


def timed(name, fns, by, reps=10):
    g = ops.GraphedCall(lambda: [f() for f in fns], device=dev)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (reps * len(fns))
    tail = f"  {by / ms / 1e6:7.0f} GB/s  frac {by / ms / 1e6 / PEAK:.3f}" if by else ""
    print(f"{name:66s} {ms * 1e3:8.1f} us{tail}", flush=True)


def token_maps(B, g, S):
    tok = torch.rand(B, g, g, device=dev, generator=gen) ** 3
    return ops.maps_from_tokens(tok, (S, S))


R = 6
for B, S, g in ((256, 336, 24), (64, 1344, 48)):
    mx, my = token_maps(B, g, S)
    for C, layout in ((1, "hwc"), (4, "hwc"), (3, "chw"), (3, "hwc")):
        shape = (B, S, S, C) if layout == "hwc" else (B, C, S, S)
        imgs = [torch.randint(0, 256, shape, device=dev, dtype=torch.uint8, generator=gen) for _ in range(R)]
        outs = [torch.empty_like(i) for i in imgs]
        timed(f"remap u8 {layout} C={C} {B} x {S}^2 (rand^3 token maps)",
              [lambda i=i, o=o: ops.remap_bilinear(i, mx, my, layout, out=o) for i, o in zip(imgs, outs)], 2 * imgs[0].numel())
        del imgs, outs
for B, S in ((256, 336), (64, 1344)):
    M = [torch.randint(0, 256, (B, S, S), device=dev, dtype=torch.uint8, generator=gen) for _ in range(4)]
    F = [m.float() / 255 for m in M[:2]]
    for tr in ("identity", "sqrt", "square", "exp", "log"):
        timed(f"maps_from_attention f32 {B} x {S}^2 {tr}", [lambda m=m: ops.maps_from_attention(m, (S, S), tr) for m in F], F[0].numel() * 4)
    for tr in ("identity", "sqrt", "log"):
        timed(f"maps_from_attention u8 {B} x {S}^2 {tr}", [lambda m=m: ops.maps_from_attention(m, (S, S), tr, 0.02, 1.0) for m in M], M[0].numel())
    del M, F
