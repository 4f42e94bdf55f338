"""GPU parity of the ragged-batch driver (BASELINE configs[3]: mixed resolutions, one launch per
stage) against the oracle image by image, and against the uniform-batch kernels."""

import numpy as np
import pytest
import torch

from gpu_util import dev, need_gpu
from oracle import numpy_path as ON

pytestmark = pytest.mark.gpu


def _tokens(n, g, seed):
    rng = np.random.default_rng(seed)
    t = rng.random((n, g, g)) ** 3
    return (t / t.sum(axis=(1, 2), keepdims=True)).astype(np.float32)


def _check(imgs, toks, out_sizes, outs, transform, max_off=1e-3):
    for im, tk, (ho, wo), o in zip(imgs, toks, out_sizes, outs):
        full = ON.upsample_tokens_nearest(tk, im.shape[0], im.shape[1])
        ref = ON.warp_image_by_attention(im, full, wo, ho, transform)
        diff = np.abs(o.cpu().numpy().astype(np.int32) - ref.astype(np.int32))
        assert diff.max() <= 1, f"{im.shape}->{(ho, wo)}: {diff.max()} LSB"
        assert (diff != 0).mean() <= max_off


@pytest.mark.parametrize("C,transform", [(3, "identity"), (3, "sqrt"), (1, "identity"), (4, "square")])
def test_ragged_vs_oracle(C, transform):
    """Sizes chosen to hit: rows that are / are not multiples of 16 bytes, one and several strips,
    strips narrower than a warp, output sizes different from the input, odd sizes."""
    need_gpu()
    from attwarp_b200 import ops
    rng = np.random.default_rng(11 + C)
    sizes = [(224, 224), (336, 336), (97, 53), (500, 333), (240, 400), (64, 1000), (701, 18), (336, 336)]
    out_sizes = [(224, 224), (500, 500), (64, 200), (500, 333), (300, 800), (64, 1000), (350, 40), (336, 336)]
    imgs = [rng.integers(0, 256, (h, w, C), dtype=np.uint8) for h, w in sizes]
    toks = _tokens(len(sizes), 24, seed=5)
    outs = ops.warp_ragged_from_tokens(dev(toks), [dev(i) for i in imgs], out_sizes, transform=transform)
    torch.cuda.synchronize()
    _check(imgs, toks, out_sizes, outs, transform)


@pytest.mark.parametrize("min_px", ["0", "24000000"])
def test_ragged_every_width_class(min_px, monkeypatch):
    """Stage 5 groups the images of a ragged batch by the consumer warps their strips need (one class per 128 output
    columns from 384 to 2048, wider images cut into strips) and by whether their rows are 4-byte aligned (direct
    stores) or not (output tiles): one image per class, both alignments, with and without the merging of small
    classes into wider ones."""
    need_gpu()
    from attwarp_b200 import ops
    monkeypatch.setenv("ATTWARP_QUAD_CLASS_MIN_PX", min_px)
    rng = np.random.default_rng(23)
    widths = [100, 384, 388, 512, 516, 640, 700, 768, 900, 1024, 1152, 1280, 1283, 1408, 1536, 1664, 1792, 1920, 2048,
              2052, 2600, 4100, 1345, 350]
    sizes = [(20 + (k * 7) % 23, max(16, w - (k % 3) * 9)) for k, w in enumerate(widths)]
    out_sizes = [(24 + (k * 5) % 19, w) for k, w in enumerate(widths)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in sizes]
    toks = _tokens(len(sizes), 8, seed=6)
    outs = ops.warp_ragged_from_tokens(dev(toks), [dev(i) for i in imgs], out_sizes)
    torch.cuda.synchronize()
    _check(imgs, toks, out_sizes, outs, "identity", max_off=5e-3)


def test_ragged_matches_uniform_batch():
    """The same images through the ragged table and through the uniform-batch entry: bit-equal."""
    need_gpu()
    from attwarp_b200 import ops
    rng = np.random.default_rng(3)
    n, H, W = 9, 336, 336
    imgs = rng.integers(0, 256, (n, H, W, 3), dtype=np.uint8)
    toks = _tokens(n, 24, seed=9)
    d_imgs = dev(imgs)
    mx, my = ops.maps_from_tokens(dev(toks), (H, W), (H, W))
    uni = ops.remap_bilinear(d_imgs, mx, my, "hwc")
    rag = ops.warp_ragged_from_tokens(dev(toks), [d_imgs[i] for i in range(n)])
    torch.cuda.synchronize()
    for i in range(n):
        assert torch.equal(uni[i], rag[i])


def test_ragged_degenerate_and_errors():
    need_gpu()
    from attwarp_b200 import ops
    from attwarp_b200._lib import AttWarpError
    assert ops.warp_ragged_from_tokens(torch.empty(0, 8, 8, device="cuda"), []) == []
    with pytest.raises(ValueError):
        ops.warp_ragged_from_tokens(torch.empty(2, 8, 8, device="cuda"), [])
    rng = np.random.default_rng(4)
    sizes = [(1, 50), (40, 40), (30, 1)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in sizes]
    toks = _tokens(3, 8, seed=2)
    out_sizes = [(4, 60), (40, 40), (33, 5)]
    outs = ops.warp_ragged_from_tokens(dev(toks), [dev(i) for i in imgs], out_sizes)
    torch.cuda.synchronize()
    _check(imgs, toks, out_sizes, outs, "identity", max_off=1.0)
    with pytest.raises(AttWarpError):
        ops.warp_ragged_from_tokens(dev(toks), [dev(np.zeros((8, 8, 2), np.uint8))] * 3)


def test_ragged_mixed_resolution_c4_sample():
    """A slice of BASELINE configs[3] (sides uniform in [224, 2048], 24x24 token maps): every
    image against the oracle, plus determinism of a second run."""
    need_gpu()
    from attwarp_b200 import ops
    rng = np.random.default_rng(1237)
    sides = rng.integers(224, 2049, size=12)
    imgs = [rng.integers(0, 256, (int(s), int(s), 3), dtype=np.uint8) for s in sides]
    toks = _tokens(len(sides), 24, seed=1237)
    d_imgs = [dev(i) for i in imgs]
    outs = ops.warp_ragged_from_tokens(dev(toks), d_imgs)
    again = ops.warp_ragged_from_tokens(dev(toks), d_imgs)
    torch.cuda.synchronize()
    sizes = [(int(s), int(s)) for s in sides]
    _check(imgs, toks, sizes, outs, "identity")
    for a, b in zip(outs, again):
        assert torch.equal(a, b)


def test_ragged_c4_nonsquare_resized():
    """BASELINE configs[3] with what real mixed-resolution batches look like: non-square images, widths that
    are not multiples of 16 (row pitches off the 16-byte phase: the per-row copy path), output sizes different
    from the input sizes (up- and down-scaling per axis) -- every image against the oracle, and the same images
    one by one through the uniform entry points (bit-equal)."""
    need_gpu()
    from attwarp_b200 import ops
    rng = np.random.default_rng(4242)
    n = 14
    hs, ws = rng.integers(224, 2049, size=n), rng.integers(224, 2049, size=n)
    ws[:4] = [333, 1001, 2047, 225]                      # 3 W mod 16 != 0
    hos = np.clip((hs * rng.uniform(0.6, 1.5, n)).astype(int), 64, 2048)
    wos = np.clip((ws * rng.uniform(0.6, 1.5, n)).astype(int), 64, 2048)
    wos[:3] = [500, 777, 1919]
    assert any((3 * int(w)) % 16 for w in ws) and any((3 * int(w)) % 16 for w in wos)
    imgs = [rng.integers(0, 256, (int(h), int(w), 3), dtype=np.uint8) for h, w in zip(hs, ws)]
    out_sizes = [(int(a), int(b)) for a, b in zip(hos, wos)]
    toks = _tokens(n, 24, seed=4242)
    d_imgs = [dev(i) for i in imgs]
    outs = ops.warp_ragged_from_tokens(dev(toks), d_imgs, out_sizes)
    torch.cuda.synchronize()
    _check(imgs, toks, out_sizes, outs, "identity")
    for i in (0, 3, n - 1):
        mx, my = ops.maps_from_tokens(dev(toks[i:i + 1]), imgs[i].shape[:2], out_sizes[i])
        one = ops.remap_bilinear(d_imgs[i][None], mx, my, "hwc")[0]
        assert torch.equal(one, outs[i])


def test_sharded_equals_unsharded_ragged():
    """multi_gpu.warp_ragged_sharded: the shards of a 2-, 3- and 8-way LPT plan (run one after the other on this
    GPU, as the ranks of a box would run them side by side) produce, image by image, exactly the bytes of the
    unsharded launch -- the per-image checksum table the multi-GPU bench gathers must not depend on the split."""
    need_gpu()
    from attwarp_b200 import multi_gpu, ops, sharding
    rng = np.random.default_rng(99)
    n = 20
    sizes = [(int(h), int(w)) for h, w in rng.integers(64, 700, (n, 2))]
    out_sizes = [(int(h), int(w)) for h, w in rng.integers(64, 700, (n, 2))]
    imgs = [dev(rng.integers(0, 256, (h, w, 3), dtype=np.uint8)) for h, w in sizes]
    toks = dev(_tokens(n, 24, seed=99))
    whole = ops.warp_ragged_from_tokens(toks, imgs, out_sizes)
    ref = [sharding.checksum64(o) for o in whole]
    for world in (2, 3, 8):
        plan = multi_gpu.plan_ragged(sizes, out_sizes, world=world)
        table = [None] * n
        for r in range(world):
            idx, outs = multi_gpu.warp_ragged_sharded(plan, toks, lambda i: imgs[i], rank=r)
            for i, o in zip(idx, outs):
                assert torch.equal(o, whole[i])
                table[i] = sharding.checksum64(o)
        assert table == ref


def test_ragged_batch_object_matches_call():
    """ops.RaggedBatch (descriptor table built once, re-used across calls) == warp_ragged_from_tokens, for two
    different token-map sets through the same object."""
    need_gpu()
    from attwarp_b200 import ops
    rng = np.random.default_rng(21)
    sizes = [(224, 224), (336, 336), (97, 53), (500, 333), (240, 400), (701, 18), (900, 1500)]
    out_sizes = [(224, 224), (500, 500), (64, 200), (500, 333), (300, 800), (350, 40), (700, 1100)]
    imgs = [dev(rng.integers(0, 256, (h, w, 3), dtype=np.uint8)) for h, w in sizes]
    batch = ops.RaggedBatch(imgs, out_sizes)
    for seed in (1, 2):
        toks = dev(_tokens(len(sizes), 24, seed=seed))
        ref = ops.warp_ragged_from_tokens(toks, imgs, out_sizes)
        got = batch.run(toks)
        torch.cuda.synchronize()
        for a, b in zip(ref, got):
            assert torch.equal(a, b)
