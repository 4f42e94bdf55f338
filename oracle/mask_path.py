"""CPU oracle for the mask post-processing between stage 1 and stage 2b of the driver flow
(TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Restates ``revise_mask`` + the mask half of ``blend_mask``
("Attention Guided Warping/attention_extraction/llava.py:207-256"):

    normalize(min) -> enhance (z-score x coe, sigmoid, clamp) -> k x k box filter with replicate padding
    -> ToPILImage (float -> uint8 by truncation of v * 255) -> PIL resize(image.size, LANCZOS) -> mode 'L'

Third-party arithmetic restated here: Pillow's 8-bit two-pass resampler (src/libImaging/Resample.c;
Pillow is unpinned by the reference, 12.2.0 in this image) -- horizontal pass then vertical pass, each
with coefficients normalised in double, rounded to 22-bit fixed point, accumulated in int32 from
1 << 21 and clipped to uint8 BETWEEN the passes.  ``tests/test_mask_path.py`` checks this restatement
bit for bit against Pillow itself and against outputs of the reference's ``blend_mask``.
"""

from __future__ import annotations

import math

import numpy as np

F32 = np.float32
PRECISION_BITS = 32 - 8 - 2


def revise_mask(patch_mask, kernel_size=3, enhance_coe=10):
    """llava.py:207-238 in float32: [gh, gw] -> [gh, gw] (values in [0, 1])."""
    m = np.asarray(patch_mask, dtype=F32)
    m = ((m - m.min()) / (m.max() - m.min())).astype(F32)                       # normalize(mat, "min")
    m = (m - m.mean(dtype=F32)).astype(F32)                                     # enhance
    std = F32(math.sqrt(float((m.astype(np.float64) ** 2).sum() / (m.size - 1))))   # torch.std: unbiased
    m = (m / std).astype(F32) * F32(enhance_coe)
    m = (F32(1) / (F32(1) + np.exp(-m.astype(F32)))).astype(F32)
    m = np.clip(m, F32(0), F32(1))
    pad = (kernel_size - 1) // 2
    p = np.pad(m, pad, mode="edge")
    w = F32(1.0) / F32(kernel_size ** 2)
    out = np.zeros_like(m)
    for dy in range(kernel_size):
        for dx in range(kernel_size):
            out = (out + w * p[dy:dy + m.shape[0], dx:dx + m.shape[1]]).astype(F32)
    return out


def to_pil_u8(mask_f32):
    """torchvision ToPILImage on a float tensor: pic.mul(255).byte() (truncation)."""
    return (np.asarray(mask_f32, dtype=F32) * F32(255)).astype(np.uint8)


def _lanczos(x):
    def sinc(t):
        if t == 0.0:
            return 1.0
        t = t * math.pi
        return math.sin(t) / t
    return sinc(x) * sinc(x / 3.0) if -3.0 <= x < 3.0 else 0.0


def lanczos_coeffs(in_size, out_size):
    """precompute_coeffs + normalize_coeffs_8bpc of Resample.c for the Lanczos filter (support 3):
    returns bounds [out, 2] = (first tap, number of taps) and fixed-point weights [out, ksize]."""
    scale = in_size / out_size
    fscale = max(scale, 1.0)
    support = 3.0 * fscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / fscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_lanczos((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _clip8(acc):
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_lanczos_u8(img, out_w, out_h):
    """PIL.Image.resize((out_w, out_h), LANCZOS) for a mode-'L' image."""
    cur = np.asarray(img, dtype=np.uint8).astype(np.int64)
    h, w = cur.shape
    if out_w != w:
        b, k = lanczos_coeffs(w, out_w)
        acc = np.zeros((h, out_w), np.int64)
        for xx in range(out_w):
            x0, n = b[xx]
            acc[:, xx] = (1 << (PRECISION_BITS - 1)) + (cur[:, x0:x0 + n] * k[xx, :n].astype(np.int64)).sum(axis=1)
        cur = _clip8(acc).astype(np.int64)
    if out_h != h:
        b, k = lanczos_coeffs(h, out_h)
        acc = np.zeros((out_h, cur.shape[1]), np.int64)
        for yy in range(out_h):
            y0, n = b[yy]
            acc[yy] = (1 << (PRECISION_BITS - 1)) + (cur[y0:y0 + n] * k[yy, :n, None].astype(np.int64)).sum(axis=0)
        cur = _clip8(acc).astype(np.int64)
    return cur.astype(np.uint8)


def mota_mask(mask_tokens, image_hw, kernel_size=3, enhance_coe=10):
    """The uint8 [H, W] mask blend_mask returns (llava.py:240-256), from a [gh, gw] token map."""
    H, W = image_hw
    return resize_lanczos_u8(to_pil_u8(revise_mask(mask_tokens, kernel_size, enhance_coe)), W, H)
