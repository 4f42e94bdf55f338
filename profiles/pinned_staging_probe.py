#!/usr/bin/env python
"""The host <-> device copies of one single-image call (0.45 MB in, 0.34 / 0.75 MB out) alone: from / to pageable
buffers directly, through pinned staging buffers, and the CPU memcpy part of the latter.  Never a bench number.

    python profiles/pinned_staging_probe.py
"""
import time, numpy as np, torch
rng = np.random.default_rng(0)
H = W = 336
for Ho in (336, 500):
    img = torch.from_numpy(rng.integers(0, 256, (H, W, 3), dtype=np.uint8))
    att = torch.from_numpy(rng.integers(0, 256, (H, W), dtype=np.uint8))
    h_out = torch.empty(Ho, Ho, 3, dtype=torch.uint8)
    p_img, p_att, p_out = img.clone().pin_memory(), att.clone().pin_memory(), torch.empty(Ho, Ho, 3, dtype=torch.uint8).pin_memory()
    d_img, d_att, d_out = img.cuda(), att.cuda(), torch.empty(Ho, Ho, 3, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    n = 1000
    def pageable():
        d_att.copy_(att); d_img.copy_(img); h_out.copy_(d_out)
    def pinned():
        p_att.copy_(att); d_att.copy_(p_att, non_blocking=True)
        p_img.copy_(img); d_img.copy_(p_img, non_blocking=True)
        p_out.copy_(d_out, non_blocking=True); torch.cuda.synchronize(); h_out.copy_(p_out)
    def memcpy_only():
        p_att.copy_(att); p_img.copy_(img); h_out.copy_(p_out)
    for name, fn in (("pageable", pageable), ("pinned staging", pinned), ("cpu memcpy only", memcpy_only)):
        for _ in range(50): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(n): fn()
        torch.cuda.synchronize(); t1 = time.perf_counter()
        print(f"out {Ho}^2: {name:18s} {(t1 - t0) / n * 1e6:7.1f} us per call")
