#!/usr/bin/env python
"""Event timings of every entry point besides the three benchmark kernels, at realistic sizes
(never a source of bench numbers; it tells which secondary kernel deserves work next).

    python profiles/kernel_survey.py
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from attwarp_b200 import checkpoint_utils as CU  # noqa: E402
from attwarp_b200 import model as M  # noqa: E402
from attwarp_b200 import new_method, ops  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)


def timed(fn, iters=8, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return min(ts), sorted(ts)[len(ts) // 2]


def report(name, fn, nbytes=None, **kw):
    best, med = timed(fn, **kw)
    extra = f"  {nbytes / best / 1e3:8.1f} GB/s (best)" if nbytes else ""
    print(f"{name:58s} best {best:9.1f} us  median {med:9.1f} us{extra}", flush=True)


# ---- stage 1 variants --------------------------------------------------------------------------------
for dt, nm in ((torch.float16, "fp16"), (torch.float32, "fp32")):
    a = torch.softmax(torch.randn(128, 32, 32, 576, device=dev, generator=g), -1).to(dt)
    out = torch.empty(128, 576, device=dev)
    report(f"aggregate [128,32,32,576] {nm}", lambda: ops.aggregate_attention(a, out=out), a.numel() * a.element_size())
    del a
# live-hook layout: [B, Hh, q, kv] fp16, last query row, per-sample token offsets
B, Hh, q, kv = 16, 32, 64, 700
att = torch.softmax(torch.randn(B, Hh, q, kv, device=dev, generator=g), -1).to(torch.float16)
view = att[:, :, -1, :].unsqueeze(1)                       # [B, 1, Hh, kv]
starts = torch.randint(1, 64, (B,), device=dev, dtype=torch.int32, generator=g)
out = torch.empty(B, 576, device=dev)
report("aggregate live hook [16,32,q=64,kv=700] fp16 + offsets", lambda: ops.aggregate_attention(view, starts, 576, out=out),
       B * Hh * 576 * 2)

# ---- stages 2-4 variants -----------------------------------------------------------------------------
tok = torch.rand(256, 24, 24, device=dev, generator=g) ** 3
report("maps_from_tokens 256 x 24^2 -> 336", lambda: ops.maps_from_tokens(tok, (336, 336)))
tok64 = torch.rand(64, 48, 48, device=dev, generator=g) ** 3
report("maps_from_tokens 64 x 48^2 -> 1344", lambda: ops.maps_from_tokens(tok64, (1344, 1344)))
att8 = torch.randint(0, 256, (256, 336, 336), device=dev, dtype=torch.uint8, generator=g)
report("maps_from_attention u8 256 x 336^2 (2 launches)", lambda: ops.maps_from_attention(att8, (336, 336)), att8.numel())
att8b = torch.randint(0, 256, (64, 1344, 1344), device=dev, dtype=torch.uint8, generator=g)
report("maps_from_attention u8 64 x 1344^2 (2 launches)", lambda: ops.maps_from_attention(att8b, (1344, 1344)), att8b.numel())
A = torch.rand(128, 1, 512, 512, device=dev, generator=g)
report("gt_marginals 128 x 512^2 f32", lambda: CU.gt_marginals(A), A.numel() * 4)
report("adaptive_avg_pool2d 128 x 512^2 -> 24^2", lambda: CU.adaptive_avg_pool2d_24(A), A.numel() * 4)
px = torch.softmax(torch.randn(128, 24, device=dev, generator=g), -1)
report("safe_softmax + mix_with_uniform [128,24] (two calls)", lambda: M.mix_with_uniform(M.safe_softmax(px), 0.1))
report("safe_softmax_mix [128,24] (one launch)", lambda: M.safe_softmax_mix(px, 0.1))
from attwarp_b200 import trainer as TR  # noqa: E402
py_ = torch.softmax(torch.randn(128, 24, device=dev, generator=g), -1)
gx_ = torch.softmax(torch.randn(128, 24, device=dev, generator=g), -1)
gy_ = torch.softmax(torch.randn(128, 24, device=dev, generator=g), -1)
report("pdf_l1_loss forward [128,24]x2 -> 512^2 (one launch)", lambda: TR.pdf_l1_loss(px, py_, gx_, gy_, (512, 512)))
pxg = px.clone().requires_grad_(True)
def _fb():
    pxg.grad = None
    TR.pdf_l1_loss(pxg, py_, gx_, gy_, (512, 512)).backward()
report("pdf_l1_loss forward + backward", _fb)
# live hook logger step (cached offsets tensor)
from attwarp_b200 import attention_extraction as AE  # noqa: E402
lg = AE.BatchMaskHookLogger(None, dev)
lg.set_batch_image_token_ranges([int(x) for x in starts.tolist()], [int(x) + 576 for x in starts.tolist()])
report("BatchMaskHookLogger._process_attention [16,32,64,700] fp16", lambda: lg._process_attention(att))
up = CU.upsample_pdf_right_inverse(px, 512)
report("upsample_pdf_right_inverse [128,24] -> 512", lambda: CU.upsample_pdf_right_inverse(px, 512))
F = CU.cdf_from_density(up.clamp_min(0))
report("cdf_from_density [128,512]", lambda: CU.cdf_from_density(up))
report("resample_cdf [128,512] -> 1344", lambda: CU.resample_cdf(F, 1344))

# ---- mask path (N2) ----------------------------------------------------------------------------------
report("mota_mask 256 x 24^2 -> 336^2 (revise + LANCZOS)", lambda: ops.mota_mask(tok, (336, 336)), 256 * 336 * 336)
report("mota_mask 64 x 24^2 -> 1344^2", lambda: ops.mota_mask(tok[:64], (1344, 1344)), 64 * 1344 * 1344)

# ---- stage 5 variants --------------------------------------------------------------------------------
mx, my = ops.maps_from_tokens(tok, (336, 336))
for C, lay, shape in ((1, "hwc", (256, 336, 336, 1)), (4, "hwc", (256, 336, 336, 4)), (3, "chw", (256, 3, 336, 336))):
    img = torch.randint(0, 256, shape, device=dev, dtype=torch.uint8, generator=g)
    o = torch.empty_like(img)
    report(f"remap u8 {lay} C={C} 256 x 336^2", lambda: ops.remap_bilinear(img, mx, my, lay, out=o), 2 * img.numel())
mx5, my5 = ops.maps_from_tokens(tok[:128], (512, 512))
for lay, shape in (("chw", (128, 3, 512, 512)), ("hwc", (128, 512, 512, 3))):
    img = torch.rand(shape, device=dev, generator=g)
    o = torch.empty_like(img)
    report(f"remap f32 {lay} 128 x 3 x 512^2", lambda: ops.remap_bilinear(img, mx5, my5, lay, out=o), 8 * img.numel())
Fx = CU.cdf_from_density(torch.rand(128, 512, device=dev, generator=g))
img = torch.rand(128, 3, 512, 512, device=dev, generator=g)
report("warp_from_cdf_torch 128 x 3 x 512^2 f32 (maps + resample)", lambda: CU.warp_from_cdf_torch(img, Fx, Fx), 8 * img.numel())
imgu = torch.randint(0, 256, (336, 336, 3), device=dev, dtype=torch.uint8, generator=g)
mx1, my1 = ops.maps_from_tokens(tok[:1], (336, 336), (500, 500))
o1 = torch.empty(1, 500, 500, 3, device=dev, dtype=torch.uint8)
report("remap u8 single 336^2 -> 500^2 (configs[0] device part)", lambda: ops.remap_bilinear(imgu[None], mx1, my1, "hwc", out=o1))

# ---- configs[0]: the NumPy drop-in, host buffers in and out, one image per call ------------------------
rng = np.random.default_rng(0)
img_h = rng.integers(0, 256, (336, 336, 3), dtype=np.uint8)
att_h = rng.integers(0, 256, (336, 336), dtype=np.uint8)
for (w, h) in ((336, 336), (500, 500)):
    for _ in range(5):
        new_method.warp_image_by_attention(img_h, att_h, w, h, transform="identity")
    t0 = time.perf_counter()
    n = 200
    for _ in range(n):
        new_method.warp_image_by_attention(img_h, att_h, w, h, transform="identity")
    dt = (time.perf_counter() - t0) / n
    print(f"{'new_method.warp_image_by_attention 336^2 -> ' + str(w) + '^2 (host in/out, wall clock)':58s} {dt * 1e6:9.1f} us per call")
try:
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from oracle import numpy_path as ON
    for _ in range(3):
        ON.warp_image_by_attention(img_h, att_h, 500, 500, "identity", remap_backend="cv2")
    t0 = time.perf_counter()
    for _ in range(20):
        ON.warp_image_by_attention(img_h, att_h, 500, 500, "identity", remap_backend="cv2")
    print(f"{'oracle port (numpy + the real cv2.remap), same call, this host':58s} {(time.perf_counter() - t0) / 20 * 1e6:9.1f} us per call")
except Exception as e:  # the oracle is test infrastructure; the survey still stands without it
    print("oracle timing skipped:", e)
