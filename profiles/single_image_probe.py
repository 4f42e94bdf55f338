#!/usr/bin/env python
"""Where the time of one new_method.warp_image_by_attention call (NumPy buffers in and out, configs[0]) goes: the whole
call, its kernels alone (graph replay on resident tensors), and the same pageable host <-> device copies alone.
Never a bench number.

    python profiles/single_image_probe.py
"""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from attwarp_b200 import new_method, ops
rng = np.random.default_rng(0)
for (H, W, Ho, Wo) in ((336, 336, 336, 336), (336, 336, 500, 500)):
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    att = rng.integers(0, 256, (H, W), dtype=np.uint8)
    for _ in range(50):
        new_method.warp_image_by_attention(img, att, Wo, Ho, transform="identity")
    t0 = time.perf_counter()
    n = 500
    for _ in range(n):
        new_method.warp_image_by_attention(img, att, Wo, Ho, transform="identity")
    t1 = time.perf_counter()
    full = (t1 - t0) / n * 1e6
    dimg = torch.from_numpy(img).cuda()[None]
    datt = torch.from_numpy(att).cuda()[None]
    out = torch.empty(1, Ho, Wo, 3, dtype=torch.uint8, device="cuda")
    def dev_step():
        mx, my = ops.maps_from_attention(datt, (Ho, Wo), "identity")
        ops.remap_bilinear(dimg, mx, my, "hwc", out=out)
    g = ops.GraphedCall(dev_step)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    dev = e0.elapsed_time(e1) / 200 * 1e3
    # copies alone: pageable numpy -> device -> pageable
    h_out = np.empty((Ho, Wo, 3), np.uint8)
    t_out = torch.from_numpy(h_out)
    t_img, t_att = torch.from_numpy(img), torch.from_numpy(att)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        datt[0].copy_(t_att); dimg[0].copy_(t_img); t_out.copy_(out[0]); 
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    cp = (t1 - t0) / n * 1e6
    print(f"{H}x{W} -> {Ho}x{Wo}: host API {full:.1f} us per call; device kernels (graph replay) {dev:.1f} us; pageable copies in + out via torch {cp:.1f} us")
