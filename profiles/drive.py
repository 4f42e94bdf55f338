#!/usr/bin/env python
"""Small kernel drivers for ncu captures (never a source of bench numbers).

    python profiles/drive.py remap  [--side 336 --batch 256 --dtype u8|f32 --layout hwc|chw --iters 3]
    python profiles/drive.py agg    [--batch 256 --dtype bf16|f16|f32]
    python profiles/drive.py maps   [--side 336 --grid 24 --batch 256]

Each runs the named stage a few times on synthetic inputs of the benchmark shapes and prints the
CUDA-event time of the last iterations, so an `ncu -k regex:...` capture is short.
"""

import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from attwarp_b200 import ops  # noqa: E402


def timed(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(iters):
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return ts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["remap", "agg", "maps", "ragged", "att"])
    ap.add_argument("--side", type=int, default=336)
    ap.add_argument("--out-side", type=int, default=0)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--grid", type=int, default=24)
    ap.add_argument("--dtype", default=None)
    ap.add_argument("--layout", default="hwc")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--C", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(0)
    B, S, G = a.batch, a.side, a.grid
    if a.what == "ragged":
        # BASELINE configs[3] slice: sides uniform in [224, 2048], 24x24 token maps, stages 2-5
        import numpy as np
        sides = np.random.default_rng(1237).integers(224, 2049, size=B)
        imgs = [torch.randint(0, 256, (int(s), int(s), a.C), device=dev, dtype=torch.uint8, generator=g) for s in sides]
        outs = [torch.empty_like(i) for i in imgs]
        tok = torch.rand(B, G, G, device=dev, generator=g) ** 3
        tok = tok / tok.sum(dim=(1, 2), keepdim=True)
        ts = timed(lambda: ops.warp_ragged_from_tokens(tok, imgs, outs=outs), a.iters)
        by = 2 * sum(i.numel() for i in imgs)
        print(f"ragged x{B} ({by / 2e6:.0f} MB in): us {[round(t, 1) for t in ts]}  GB/s {[round(by / t / 1e3, 1) for t in ts]}")
        return
    So = a.out_side or S
    if a.what == "att":
        # stage 2b of the NumPy path: marginals of a materialised attention map [B,H,W] (uint8 or float32)
        if (a.dtype or "u8") == "u8":
            sets = [torch.randint(0, 256, (B, S, S), device=dev, dtype=torch.uint8, generator=g) for _ in range(3)]
        else:
            sets = [torch.rand(B, S, S, device=dev, generator=g) for _ in range(3)]
        k = [0]

        def run():
            ops.maps_from_attention(sets[k[0] % 3], (So, So))
            k[0] += 1
        ts = timed(run, a.iters)
        by = sets[0].numel() * sets[0].element_size()
        print(f"maps_from_attention {a.dtype or 'u8'} {S}^2 x{B}: us {[round(t, 1) for t in ts]}  GB/s {[round(by / t / 1e3, 1) for t in ts]}")
        return
    if a.what in ("remap", "maps"):
        tok = torch.rand(B, G, G, device=dev, generator=g) ** 3
        tok = tok / tok.sum(dim=(1, 2), keepdim=True)
        if a.what == "maps":
            ts = timed(lambda: ops.maps_from_tokens(tok, (S, S), (So, So)), a.iters)
            print("maps_from_tokens us:", ts)
            return
        mx, my = ops.maps_from_tokens(tok, (S, S), (So, So))
        dt = a.dtype or "u8"
        shape = (B, S, S, a.C) if a.layout == "hwc" else (B, a.C, S, S)
        if dt == "u8":
            sets = [torch.randint(0, 256, shape, device=dev, dtype=torch.uint8, generator=g) for _ in range(3)]
        else:
            sets = [torch.rand(shape, device=dev, generator=g) for _ in range(3)]
        oshape = (B, So, So, a.C) if a.layout == "hwc" else (B, a.C, So, So)
        out = torch.empty(oshape, device=dev, dtype=sets[0].dtype)
        k = [0]

        def run():
            ops.remap_bilinear(sets[k[0] % 3], mx, my, a.layout, out=out)
            k[0] += 1
        ts = timed(run, a.iters)
        by = (sets[0].numel() + out.numel()) * sets[0].element_size()
        print(f"remap {dt} {a.layout} {S}->{So} x{B}: us {ts}  GB/s {[round(by / t / 1e3, 1) for t in ts]}")
    else:
        dt = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}[a.dtype or "bf16"]
        T = G * G
        sets = [torch.softmax(torch.randn(B, 32, 32, T, device=dev, generator=g), -1).to(dt) for _ in range(2)]
        out = torch.empty(B, T, device=dev)
        k = [0]

        def run():
            ops.aggregate_attention(sets[k[0] % 2], out=out)
            k[0] += 1
        ts = timed(run, a.iters)
        by = sets[0].numel() * sets[0].element_size()
        print(f"aggregate {a.dtype}: us {ts}  GB/s {[round(by / t / 1e3, 1) for t in ts]}")


if __name__ == "__main__":
    main()
