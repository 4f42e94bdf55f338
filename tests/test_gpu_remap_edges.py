"""GPU parity, stage 5 edge cases: the streaming uint8 resample kernel against the oracle's restatement
of cv2.remap(INTER_LINEAR, BORDER_REPLICATE) (itself pinned to the real cv2 in
test_oracle_vs_golden.py).  Bit-exact (0 LSB) given identical float32 maps.

Covers what the attention-derived maps of the benchmark never produce but the C ABI accepts:
non-monotone maps (per-row path), strong minification (source rows skipped / pass splitting /
direct-row path), strong magnification (many output rows per source row pair), constant and
out-of-range maps (border replicate on every tap), degenerate 1-pixel axes, odd pitches that make
every row start at a different 16-byte phase."""

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
        assert np.array_equal(out[b], ref), f"image {b}: {np.abs(out[b].astype(int) - ref.astype(int)).max()} LSB"


@pytest.mark.parametrize("layout", ["hwc", "chw"])
@pytest.mark.parametrize("C", [1, 3, 4])
def test_unsorted_maps(layout, C):
    need_gpu()
    rng = np.random.default_rng(7 + C)
    B, H, W, Ho, Wo = 2, 61, 83, 70, 97
    img = rng.integers(0, 256, (B, H, W, C), dtype=np.uint8)
    mx = (rng.random((B, Wo)) * (W + 6) - 3).astype(np.float32)
    my = (rng.random((B, Ho)) * (H + 6) - 3).astype(np.float32)
    _check(img, mx, my, layout)


@pytest.mark.parametrize("C", [1, 3, 4])
def test_strong_minification(C):
    """2048 -> 40 rows/cols: source rows between taps are skipped, spans exceed the arena."""
    need_gpu()
    rng = np.random.default_rng(17 + C)
    B, H, W, Ho, Wo = 1, 1500, 1700, 40, 45
    img = rng.integers(0, 256, (B, H, W, C), dtype=np.uint8)
    mx = np.sort(rng.random((B, Wo)) * W, axis=1).astype(np.float32)
    my = np.sort(rng.random((B, Ho)) * H, axis=1).astype(np.float32)
    _check(img, mx, my)
    # wide, strongly minified rows with an odd width: the rows whose source span does not fit a stage are gathered
    # pixel by pixel by several warps (each skips the halo column it shares with its neighbour)
    if C == 3:
        mx3 = np.sort(rng.random((B, 301)) * W, axis=1).astype(np.float32)
        my3 = np.sort(rng.random((B, 33)) * H, axis=1).astype(np.float32)
        _check(img, mx3, my3)
    # moderate vertical minification only (3.7x): rows are skipped but the pass still fits
    my2 = (np.arange(400, dtype=np.float32) * 3.7)[None]
    mx2 = (np.arange(500, dtype=np.float32) * 1.01 + 0.3)[None]
    _check(img, mx2, my2)


@pytest.mark.parametrize("C", [1, 3])
def test_strong_magnification(C):
    """12 source pixels stretched over 700 outputs: many output rows share one slot pair."""
    need_gpu()
    rng = np.random.default_rng(27 + C)
    B, H, W, Ho, Wo = 2, 12, 13, 700, 650
    img = rng.integers(0, 256, (B, H, W, C), dtype=np.uint8)
    mx = np.tile(np.linspace(-0.5, W - 0.4, Wo, dtype=np.float32), (B, 1))
    my = np.tile(np.linspace(-0.5, H - 0.4, Ho, dtype=np.float32), (B, 1))
    _check(img, mx, my)


def test_constant_and_out_of_range_maps():
    need_gpu()
    rng = np.random.default_rng(3)
    B, H, W, Ho, Wo = 3, 33, 47, 50, 60
    img = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)
    mx = np.stack([np.full(Wo, -7.3), np.full(Wo, W + 100.0), np.full(Wo, 11.49)]).astype(np.float32)
    my = np.stack([np.full(Ho, H + 9.0), np.full(Ho, -1e6), np.full(Ho, 5.5)]).astype(np.float32)
    _check(img, mx, my)
    # decreasing maps (a flip)
    mx = np.tile(np.linspace(W - 1, 0, Wo, dtype=np.float32), (B, 1))
    my = np.tile(np.linspace(H - 1, 0, Ho, dtype=np.float32), (B, 1))
    _check(img, mx, my)


@pytest.mark.parametrize("H,W", [(1, 50), (50, 1), (1, 1), (2, 2)])
def test_degenerate_axes(H, W):
    need_gpu()
    rng = np.random.default_rng(H * 100 + W)
    img = rng.integers(0, 256, (2, H, W, 3), dtype=np.uint8)
    mx = np.sort(rng.random((2, 37)) * (W + 2) - 1, axis=1).astype(np.float32)
    my = np.sort(rng.random((2, 29)) * (H + 2) - 1, axis=1).astype(np.float32)
    _check(img, mx, my)
    _check(img, mx, my, "chw")


@pytest.mark.parametrize("W,Wo", [(333, 335), (500, 501), (1021, 777)])
def test_odd_pitches(W, Wo):
    """Row pitches that are not multiples of 4 / 16 bytes: every staged row has its own phase."""
    need_gpu()
    rng = np.random.default_rng(W)
    H, Ho = 77, 91
    img = rng.integers(0, 256, (2, H, W, 3), dtype=np.uint8)
    mx = np.sort(rng.random((2, Wo)) * W, axis=1).astype(np.float32)
    my = np.sort(rng.random((2, Ho)) * H, axis=1).astype(np.float32)
    _check(img, mx, my)
    # a view with a non-zero storage offset: the image base itself is misaligned
    from attwarp_b200 import ops
    import torch
    flat = torch.zeros(img[0].size + 5, dtype=torch.uint8, device="cuda")
    flat[5:] = torch.from_numpy(img[0]).cuda().reshape(-1)
    view = flat[5:].view(1, H, W, 3)
    out = ops.remap_bilinear(view, dev(mx[:1]), dev(my[:1]), "hwc")[0].cpu().numpy()
    assert np.array_equal(out, ON.remap(img[0], mx[0], my[0]))


@pytest.mark.parametrize("Wo", [1, 2, 5, 127, 128, 129, 335, 336, 385, 500, 1000, 1347, 2047, 2050])
@pytest.mark.parametrize("policy", ["0", "2"])
def test_destination_at_every_alignment(Wo, policy, monkeypatch):
    """3-channel rows written straight from the consumer warps: a destination whose first byte sits 0..3 bytes past
    a 4-byte boundary and whose row pitch 3 * Wo walks through every alignment.  Guard bytes around the destination
    must stay untouched (the partial first / last words of a warp's block are stored byte by byte)."""
    need_gpu()
    from attwarp_b200 import ops
    monkeypatch.setenv("ATTWARP_QUAD_MAP", policy)
    rng = np.random.default_rng(Wo)
    B, H, W, Ho = 2, 37, max(2, Wo - 3), 41
    img = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)
    mx = np.sort(rng.random((B, Wo)) * W, axis=1).astype(np.float32)
    my = np.sort(rng.random((B, Ho)) * H, axis=1).astype(np.float32)
    refs = [ON.remap(img[b], mx[b], my[b]) for b in range(B)]
    n = B * Ho * Wo * 3
    for off in range(4):
        flat = torch.full((n + 64,), 0xA5, dtype=torch.uint8, device="cuda")
        out = flat[16 + off:16 + off + n].view(B, Ho, Wo, 3)
        ops.remap_bilinear(dev(img), dev(mx), dev(my), "hwc", out=out)
        torch.cuda.synchronize()
        got = flat.cpu().numpy()
        assert (got[:16 + off] == 0xA5).all() and (got[16 + off + n:] == 0xA5).all(), f"offset {off}: guard bytes overwritten"
        o = got[16 + off:16 + off + n].reshape(B, Ho, Wo, 3)
        for b in range(B):
            assert np.array_equal(o[b], hwc(refs[b])), f"offset {off}, image {b}"


@pytest.mark.parametrize("H,W,Ho,Wo", [(336, 336, 336, 336), (336, 336, 500, 500), (1344, 1344, 1344, 1344),
                                       (301, 224, 500, 500), (500, 333, 400, 700), (97, 53, 64, 200)])
def test_quad_kernel_mappings_agree(H, W, Ho, Wo, monkeypatch):
    """The 3-channel kernel (remap_quad.cu) picks, per warp and strip, between two thread <-> pixel mappings (QUAD:
    four adjacent columns per thread; LANE: columns 32 apart + an in-warp transpose) and between fixed-shift and
    per-slot addressing; every combination, and the round-1 kernel, must give the same bytes as the direct
    (one thread per pixel) kernel on smooth and on strongly non-uniform maps."""
    need_gpu()
    from attwarp_b200 import ops
    rng = np.random.default_rng(H * 7 + Wo)
    B = 3
    img = dev(rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8))
    maps = []
    for kind in ("near_identity", "rand3", "steps"):
        if kind == "near_identity":
            tok = 1.0 + 0.02 * rng.standard_normal((B, 24, 24))
        elif kind == "rand3":
            tok = rng.random((B, 24, 24)) ** 3
        else:
            tok = np.where(rng.random((B, 24, 24)) < 0.3, 20.0, 0.05)
        tok = (tok / tok.sum(axis=(1, 2), keepdims=True)).astype(np.float32)
        maps.append(ops.maps_from_tokens(dev(tok), (H, W), (Ho, Wo)))
    for mx, my in maps:
        monkeypatch.setenv("ATTWARP_REMAP", "direct")
        # the direct kernel is selected once per process (static), so the reference here is the oracle-checked
        # default path with the QUAD-only policy; policies are read at every launch
        monkeypatch.delenv("ATTWARP_REMAP")
        outs = {}
        for policy in ("1", "2", "0"):
            monkeypatch.setenv("ATTWARP_QUAD_MAP", policy)
            outs[policy] = ops.remap_bilinear(img, mx, my, "hwc")
        torch.cuda.synchronize()
        assert torch.equal(outs["1"], outs["2"]) and torch.equal(outs["1"], outs["0"])
        # and against the oracle's integer formula on one image
        from oracle import numpy_path as ON
        ref = ON.remap_u8(img[0].cpu().numpy(), mx[0].cpu().numpy(), my[0].cpu().numpy())
        assert np.array_equal(outs["0"][0].cpu().numpy(), ref)


def _maps_family(rng, kind, n_out, n_in, B):
    """Separable map rows of one family: attention-like (monotone, near the identity), random monotone, unsorted,
    strongly shrinking, constant / out of range."""
    rows = []
    for _ in range(B):
        if kind == "near_identity":
            m = np.linspace(0, n_in - 1, n_out) + rng.normal(0, 0.4, n_out)
        elif kind == "monotone":
            m = np.sort(rng.random(n_out) * n_in)
        elif kind == "unsorted":
            m = rng.random(n_out) * (n_in + 4) - 2
        elif kind == "shrink":
            m = np.sort(rng.random(n_out)) ** 3 * n_in
        else:
            m = np.concatenate([np.full(n_out // 2, -3.0), np.full(n_out - n_out // 2, n_in + 2.5)])
        rows.append(m.astype(np.float32))
    return np.stack(rows)


@pytest.mark.parametrize("kind", ["near_identity", "monotone", "unsorted", "shrink", "out_of_range"])
@pytest.mark.parametrize("C,layout", [(1, "hwc"), (4, "hwc"), (3, "chw"), (2, "chw"), (4, "chw")])
@pytest.mark.parametrize("H,W,Ho,Wo", [(61, 64, 70, 64), (45, 336, 129, 500), (33, 203, 40, 131), (130, 8, 67, 4)])
def test_walk_kernel_formats(kind, C, layout, H, W, Ho, Wo):
    """Grey, 4-channel and planar uint8 images (remap_stream.cu; written for the column-walking kernel of
    profiles/r06_walk_kernel.md, which passed it too): bit-equal to the cv2.remap restatement for every map family,
    with word-aligned and odd pitches on both sides, tall outputs and border rows."""
    need_gpu()
    rng = np.random.default_rng(H * 131 + W * 7 + C)
    B = 2
    img = rng.integers(0, 256, (B, H, W, C), dtype=np.uint8)
    _check(img, _maps_family(rng, kind, Wo, W, B), _maps_family(rng, kind, Ho, H, B), layout)


@pytest.mark.parametrize("C,layout", [(1, "hwc"), (4, "hwc"), (3, "chw")])
@pytest.mark.parametrize("src_off,dst_off", [(0, 1), (1, 0), (2, 3), (4, 4)])
def test_walk_kernel_misaligned_bases(C, layout, src_off, dst_off):
    """Views with non-zero storage offsets: source and destination bases at every phase of a 32-bit word."""
    need_gpu()
    from attwarp_b200 import ops
    rng = np.random.default_rng(C * 17 + src_off * 5 + dst_off)
    B, H, W, Ho, Wo = 2, 50, 64, 70, 72
    img = rng.integers(0, 256, (B, H, W, C), dtype=np.uint8)
    mx = _maps_family(rng, "monotone", Wo, W, B)
    my = _maps_family(rng, "near_identity", Ho, H, B)
    arr = img if layout == "hwc" else np.ascontiguousarray(np.transpose(img, (0, 3, 1, 2)))
    flat = torch.zeros(arr.size + src_off, dtype=torch.uint8, device="cuda")
    flat[src_off:] = torch.from_numpy(arr).cuda().reshape(-1)
    src = flat[src_off:].view(arr.shape)
    oshape = (B, Ho, Wo, C) if layout == "hwc" else (B, C, Ho, Wo)
    oflat = torch.full((int(np.prod(oshape)) + dst_off + 8,), 0xAB, dtype=torch.uint8, device="cuda")
    out = oflat[dst_off:dst_off + int(np.prod(oshape))].view(oshape)
    ops.remap_bilinear(src, dev(mx), dev(my), layout, out=out)
    got = out.cpu().numpy()
    if layout == "chw":
        got = np.transpose(got, (0, 2, 3, 1))
    for b in range(B):
        assert np.array_equal(got[b], hwc(ON.remap(img[b] if C > 1 else img[b][..., 0], mx[b], my[b])))
    # nothing written outside the destination view
    assert bool((oflat[:dst_off] == 0xAB).all()) and bool((oflat[dst_off + int(np.prod(oshape)):] == 0xAB).all())
