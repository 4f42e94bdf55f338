"""Stage 5 for float32 images (remap_f32_stream.cu; the torch path of the reference resamples float images with
cv2.remap, checkpoint_utils.py:195-198): bit-equal to the oracle's restatement of OpenCV's float path (which
tests/test_oracle_vs_golden.py holds bit-equal to the real cv2.remap) on the same edge cases as the uint8 suite:
unsorted / decreasing / constant / out-of-range maps, strong minification and magnification, 1-pixel axes, odd
sizes, interleaved and planar layouts, several strips, and the BASELINE configs[4] shape."""

import numpy as np
import pytest
import torch

from gpu_util import dev, hwc, need_gpu
from oracle import numpy_path as ON

pytestmark = pytest.mark.gpu


def _check(img, mx, my, layout="hwc"):
    from attwarp_b200 import ops
    B = img.shape[0]
    src = dev(img if layout == "hwc" else np.ascontiguousarray(np.transpose(img, (0, 3, 1, 2))))
    out = ops.remap_bilinear(src, dev(mx), dev(my), layout).cpu().numpy()
    if layout == "chw":
        out = np.transpose(out, (0, 2, 3, 1))
    for b in range(B):
        ref = hwc(ON.remap(img[b], mx[b], my[b]))
        assert out[b].dtype == np.float32 and np.array_equal(out[b], ref), \
            f"image {b}: max abs diff {np.abs(out[b].astype(np.float64) - ref).max():.3e}"


def _img(rng, shape):
    return (rng.random(shape) * 2 - 0.5).astype(np.float32)


@pytest.mark.parametrize("layout", ["hwc", "chw"])
@pytest.mark.parametrize("C", [1, 3, 4])
def test_f32_unsorted_maps(layout, C):
    need_gpu()
    rng = np.random.default_rng(70 + C)
    B, H, W, Ho, Wo = 2, 61, 83, 70, 97
    _check(_img(rng, (B, H, W, C)), (rng.random((B, Wo)) * (W + 6) - 3).astype(np.float32),
           (rng.random((B, Ho)) * (H + 6) - 3).astype(np.float32), layout)


@pytest.mark.parametrize("layout,C", [("chw", 3), ("hwc", 3), ("hwc", 1)])
def test_f32_minification_and_magnification(layout, C):
    need_gpu()
    rng = np.random.default_rng(170 + C)
    img = _img(rng, (1, 900, 1100, C))
    _check(img, np.sort(rng.random((1, 45)) * 1100, axis=1).astype(np.float32),
           np.sort(rng.random((1, 40)) * 900, axis=1).astype(np.float32), layout)
    _check(img, (np.arange(500, dtype=np.float32) * 1.01 + 0.3)[None], (np.arange(240, dtype=np.float32) * 3.7)[None], layout)
    small = _img(rng, (2, 12, 13, C))
    _check(small, np.tile(np.linspace(-0.5, 12.6, 650, dtype=np.float32), (2, 1)),
           np.tile(np.linspace(-0.5, 11.6, 700, dtype=np.float32), (2, 1)), layout)


def test_f32_constant_decreasing_and_out_of_range_maps():
    need_gpu()
    rng = np.random.default_rng(31)
    B, H, W, Ho, Wo = 3, 33, 47, 50, 60
    img = _img(rng, (B, H, W, 3))
    mx = np.stack([np.full(Wo, -7.3), np.full(Wo, W + 100.0), np.full(Wo, 11.49)]).astype(np.float32)
    my = np.stack([np.full(Ho, H + 9.0), np.full(Ho, -1e6), np.full(Ho, 5.5)]).astype(np.float32)
    _check(img, mx, my)
    _check(img, mx, my, "chw")
    mx = np.tile(np.linspace(W - 1, 0, Wo, dtype=np.float32), (B, 1))
    my = np.tile(np.linspace(H - 1, 0, Ho, dtype=np.float32), (B, 1))
    _check(img, mx, my, "chw")
    # up-and-down rows around the top and bottom borders (both taps on one source row, lower tap going back)
    my = np.tile(np.concatenate([np.linspace(1.2, -2.0, 20), np.linspace(-2.0, H + 3, Ho - 40), np.linspace(H + 3, H - 2.5, 20)])
                 .astype(np.float32), (B, 1))
    _check(img, np.tile(np.linspace(0, W, Wo, dtype=np.float32), (B, 1)), my, "chw")


@pytest.mark.parametrize("H,W", [(1, 50), (50, 1), (1, 1), (2, 2)])
def test_f32_degenerate_axes(H, W):
    need_gpu()
    rng = np.random.default_rng(H * 100 + W)
    img = _img(rng, (2, H, W, 3))
    mx = np.sort(rng.random((2, 37)) * (W + 2) - 1, axis=1).astype(np.float32)
    my = np.sort(rng.random((2, 29)) * (H + 2) - 1, axis=1).astype(np.float32)
    _check(img, mx, my)
    _check(img, mx, my, "chw")


@pytest.mark.parametrize("W,Wo,layout", [(333, 335, "hwc"), (501, 500, "chw"), (1021, 1777, "chw"), (700, 2500, "hwc")])
def test_f32_odd_sizes_and_strips(W, Wo, layout):
    """Row pitches off the 16-byte phase, outputs wider than one strip (per-row copies, several strips)."""
    need_gpu()
    rng = np.random.default_rng(W)
    H, Ho = 77, 91
    img = _img(rng, (2, H, W, 3))
    _check(img, np.sort(rng.random((2, Wo)) * W, axis=1).astype(np.float32),
           np.sort(rng.random((2, Ho)) * H, axis=1).astype(np.float32), layout)


def test_f32_c5_shape_both_map_families():
    """BASELINE configs[4] shape (3 x 512^2 float32 planes) with near-identity PDF maps and with rand^3 token maps;
    every image of a 16-image batch against the oracle, plus the round-1 kernel's result (bit-equal)."""
    need_gpu()
    from attwarp_b200 import ops
    rng = np.random.default_rng(512)
    B = 16
    img = rng.random((B, 3, 512, 512)).astype(np.float32)
    d_img = dev(img)
    tok_a = 1.0 + 0.05 * rng.standard_normal((B, 24, 24))
    tok_b = rng.random((B, 24, 24)) ** 3
    for tok in (tok_a, tok_b):
        tok = (tok / tok.sum(axis=(1, 2), keepdims=True)).astype(np.float32)
        mx, my = ops.maps_from_tokens(dev(tok), (512, 512))
        out = ops.remap_bilinear(d_img, mx, my, "chw").cpu().numpy()
        mxh, myh = mx.cpu().numpy(), my.cpu().numpy()
        for b in range(B):
            ref = ON.remap(np.ascontiguousarray(np.transpose(img[b], (1, 2, 0))), mxh[b], myh[b])
            assert np.array_equal(np.transpose(out[b], (1, 2, 0)), ref), b
