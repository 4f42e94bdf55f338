#!/usr/bin/env python
"""Device time of the secondary kernels of the path (stage 2 helpers, the driver-flow mask), each replayed back to
back from a CUDA graph over rotating inputs (eager single calls, as in kernel_survey.py, include 10-25 us of launch
overhead).  Never a bench number.

    python profiles/row_kernels_probe.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from attwarp_b200 import checkpoint_utils as CU, ops  # noqa: E402

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


R = 4
A = [torch.rand(128, 1, 512, 512, device=dev, generator=gen) for _ in range(R)]
timed("gt_marginals 128 x 512^2 f32 (marginals + finish)", [lambda a=a: CU.gt_marginals(a) for a in A], A[0].numel() * 4)
timed("adaptive_avg_pool2d 128 x 512^2 -> 24^2", [lambda a=a: CU.adaptive_avg_pool2d_24(a) for a in A], A[0].numel() * 4)
del A
for B, S in ((256, 336), (64, 1344)):
    M = [torch.randint(0, 256, (B, S, S), device=dev, dtype=torch.uint8, generator=gen) for _ in range(R)]
    timed(f"maps_from_attention u8 {B} x {S}^2 (marginals + finish)", [lambda m=m: ops.maps_from_attention(m, (S, S), "identity") for m in M], M[0].numel())
    F = [m.float() for m in M[:2]]
    timed(f"maps_from_attention f32 {B} x {S}^2 sqrt", [lambda m=m: ops.maps_from_attention(m, (S, S), "sqrt") for m in F], F[0].numel() * 4)
    del M, F
    T = [torch.rand(B, 24, 24, device=dev, generator=gen) for _ in range(R)]
    timed(f"mota_mask {B} x 24^2 -> {S}^2 (revise + LANCZOS, mask written)", [lambda t=t: ops.mota_mask(t, (S, S)) for t in T], B * S * S)
    timed(f"mota_mask + maps_from_attention {B} x {S}^2 (4 launches)", [lambda t=t: ops.maps_from_attention(ops.mota_mask(t, (S, S)), (S, S), "identity") for t in T], 2 * B * S * S)
    timed(f"maps_from_mota_tokens {B} x {S}^2 (fused, mask never written)", [lambda t=t: ops.maps_from_mota_tokens(t, (S, S)) for t in T], 0)
    timed(f"maps_from_tokens {B} x 24^2 -> {S}", [lambda t=t: ops.maps_from_tokens(t, (S, S)) for t in T], 0)
