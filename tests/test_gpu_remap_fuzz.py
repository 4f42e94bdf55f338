"""GPU parity, stage 5: seeded random shapes, alignments and maps through the 3-channel uint8 kernel (uniform batches
and ragged batches) against the oracle's restatement of cv2.remap.  Bit-exact.  ATTWARP_FUZZ_CASES raises the number of
cases (default 48: a few seconds)."""

import os

import numpy as np
import pytest
import torch

from gpu_util import dev, hwc, need_gpu
from oracle import numpy_path as ON

pytestmark = pytest.mark.gpu

N_CASES = int(os.environ.get("ATTWARP_FUZZ_CASES", "48"))


def _side(rng, big):
    # mostly small, sometimes around the warp / strip boundaries of the kernel, sometimes large
    r = rng.random()
    if r < 0.45:
        return int(rng.integers(2, 200))
    if r < 0.8:
        edge = int(rng.choice([127, 128, 254, 256, 381, 384, 508, 512, 635, 1016, 1024, 1270, 1397, 1408, 2032, 2048]))
        return max(2, edge + int(rng.integers(-3, 4)))
    return int(rng.integers(200, max(big, 201)))


def _map(rng, n_out, n_in, kind):
    if kind == 0:                                   # monotone, covers the image
        m = np.sort(rng.random(n_out) * n_in)
    elif kind == 1:                                 # near identity with jitter
        m = np.linspace(0, n_in - 1, n_out) + rng.normal(0, 0.7, n_out)
    elif kind == 2:                                 # arbitrary, partly outside the image
        m = rng.random(n_out) * (n_in + 8) - 4
    else:                                           # piecewise: flat runs and jumps
        m = np.repeat(rng.random(max(1, n_out // 37 + 1)) * n_in, 37)[:n_out]
    return m.astype(np.float32)


@pytest.mark.parametrize("case", range(N_CASES))
def test_uniform_batch_random(case):
    need_gpu()
    from attwarp_b200 import ops
    rng = np.random.default_rng(9000 + case)
    B = int(rng.integers(1, 4))
    H, W, Ho, Wo = _side(rng, 700), _side(rng, 2300), _side(rng, 500), _side(rng, 2300)
    img = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)
    kx, ky = int(rng.integers(0, 4)), int(rng.integers(0, 4))
    mx = np.stack([_map(rng, Wo, W, kx) for _ in range(B)])
    my = np.stack([_map(rng, Ho, H, ky) for _ in range(B)])
    off = int(rng.integers(0, 4))
    n = B * Ho * Wo * 3
    flat = torch.full((n + 32,), 0x5A, dtype=torch.uint8, device="cuda")
    out = flat[16 + off:16 + off + n].view(B, Ho, Wo, 3)
    ops.remap_bilinear(dev(img), dev(mx), dev(my), "hwc", out=out)
    torch.cuda.synchronize()
    got = flat.cpu().numpy()
    assert (got[:16 + off] == 0x5A).all() and (got[16 + off + n:] == 0x5A).all(), "guard bytes overwritten"
    o = got[16 + off:16 + off + n].reshape(B, Ho, Wo, 3)
    for b in range(B):
        ref = hwc(ON.remap(img[b], mx[b], my[b]))
        assert np.array_equal(o[b], ref), (case, (H, W, Ho, Wo), off, (kx, ky), b, int(np.abs(o[b].astype(int) - ref.astype(int)).max()))


@pytest.mark.parametrize("case", range(max(4, N_CASES // 8)))
def test_ragged_batch_random(case):
    need_gpu()
    from attwarp_b200 import ops
    rng = np.random.default_rng(9500 + case)
    n = int(rng.integers(3, 14))
    sizes = [(_side(rng, 300), _side(rng, 2300)) for _ in range(n)]
    out_sizes = [(_side(rng, 200), _side(rng, 2300)) for _ in range(n)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in sizes]
    tok = rng.random((n, 8, 8)) ** 2
    tok = (tok / tok.sum(axis=(1, 2), keepdims=True)).astype(np.float32)
    outs = ops.warp_ragged_from_tokens(dev(tok), [dev(i) for i in imgs], out_sizes)
    torch.cuda.synchronize()
    for im, tk, (ho, wo), o in zip(imgs, tok, out_sizes, outs):
        full = ON.upsample_tokens_nearest(tk, im.shape[0], im.shape[1])
        ref = ON.warp_image_by_attention(im, full, wo, ho, "identity")
        d = np.abs(o.cpu().numpy().astype(np.int32) - ref.astype(np.int32))
        assert d.max() <= 1 and (d != 0).mean() <= 5e-3, (case, im.shape, (ho, wo), int(d.max()))


_FORMATS = [("u8", "hwc", 1), ("u8", "hwc", 4), ("u8", "chw", 3), ("u8", "chw", 1), ("f32", "chw", 3), ("f32", "hwc", 3),
            ("f32", "hwc", 1), ("f32", "chw", 4)]


@pytest.mark.parametrize("case", range(max(8, N_CASES // 2)))
def test_other_formats_random(case):
    """The same random shapes through the kernels of the other formats: float32 (stream kernel and its fallbacks),
    uint8 with 1 or 4 channels and planar uint8."""
    need_gpu()
    from attwarp_b200 import ops
    rng = np.random.default_rng(9800 + case)
    dt, layout, C = _FORMATS[case % len(_FORMATS)]
    B = int(rng.integers(1, 3))
    H, W, Ho, Wo = _side(rng, 500), _side(rng, 1700), _side(rng, 400), _side(rng, 1700)
    if dt == "u8":
        img = rng.integers(0, 256, (B, H, W, C), dtype=np.uint8)
    else:
        img = (rng.random((B, H, W, C)) * 2 - 0.5).astype(np.float32)
    kx, ky = int(rng.integers(0, 4)), int(rng.integers(0, 4))
    mx = np.stack([_map(rng, Wo, W, kx) for _ in range(B)])
    my = np.stack([_map(rng, Ho, H, ky) for _ in range(B)])
    src = dev(img if layout == "hwc" else np.ascontiguousarray(np.transpose(img, (0, 3, 1, 2))))
    out = ops.remap_bilinear(src, dev(mx), dev(my), layout).cpu().numpy()
    if layout == "chw":
        out = np.transpose(out, (0, 2, 3, 1))
    for b in range(B):
        ref = hwc(ON.remap(img[b] if C > 1 else img[b][..., 0], mx[b], my[b]))
        assert np.array_equal(out[b], ref), (case, dt, layout, C, (H, W, Ho, Wo), (kx, ky), b)
