"""GPU parity of the fused batch driver (stages 1-5, BASELINE configs[1]/[2] shapes) and
size-independent properties at full benchmark sizes."""

import numpy as np
import pytest
import torch

from gpu_util import dev, need_gpu, rel_err
from oracle import aggregate as OA
from oracle import numpy_path as ON

pytestmark = pytest.mark.gpu


def _attention(B, L, Hh, T, dtype, seed):
    gen = torch.Generator().manual_seed(seed)
    a = torch.softmax(torch.randn(B, L, Hh, T + 124, generator=gen) * 2, dim=-1)[..., :T]
    return a.contiguous().to(dtype)


@pytest.mark.parametrize("B,L,Hh,gh,H,Ho,layout,transform", [
    (6, 32, 32, 24, 336, 336, "hwc", "identity"),
    (3, 4, 8, 24, 336, 500, "hwc", "sqrt"),
    (2, 2, 4, 48, 1344, 1344, "hwc", "identity"),
    (3, 4, 8, 24, 336, 336, "chw", "identity"),
])
def test_fused_batch_vs_oracle(B, L, Hh, gh, H, Ho, layout, transform):
    need_gpu()
    from attwarp_b200 import ops
    rng = np.random.default_rng(B * 7 + H)
    attn = _attention(B, L, Hh, gh * gh, torch.bfloat16, seed=H + B)
    imgs = rng.integers(0, 256, (B, H, H, 3), dtype=np.uint8)
    src = dev(imgs if layout == "hwc" else np.transpose(imgs, (0, 3, 1, 2)))
    out, tok, mx, my = ops.warp_from_attention_tokens(attn.cuda(), src, (gh, gh), (Ho, Ho), layout,
                                                      transform=transform, return_aux=True)
    out = out.cpu().numpy()
    if layout == "chw":
        out = np.transpose(out, (0, 2, 3, 1))
    tok = tok.cpu().numpy()
    tok_ref = OA.aggregate_attention(attn.float().numpy())
    assert rel_err(tok.reshape(B, -1), tok_ref, floor=1e-12) <= 1e-5
    for b in range(B):
        # stage-wise: stages 2-5 fed with the GPU's own token map (fp32), float64 oracle after
        full = ON.upsample_tokens_nearest(tok[b], H, H)
        ref, rx, ry = ON.warp_image_by_attention(imgs[b], full, Ho, Ho, transform, return_maps=True)
        assert np.abs(mx[b].cpu().numpy() - rx).max() <= 1e-4
        assert np.abs(my[b].cpu().numpy() - ry).max() <= 1e-4
        diff = np.abs(out[b].astype(np.int32) - ref.astype(np.int32))
        assert diff.max() <= 1 and (diff != 0).mean() <= 1e-3


def _check_all_images(out, imgs, tok, H):
    """All images of a full-size batch against the oracle; reports the worst image."""
    out_h, imgs_h, tok_h = out.cpu().numpy(), imgs.cpu().numpy(), tok.cpu().numpy()
    worst, n_off = 0, 0
    for b in range(out_h.shape[0]):
        full = ON.upsample_tokens_nearest(tok_h[b].reshape(tok_h.shape[-2:]) if tok_h.ndim == 3 else tok_h[b], H, H)
        ref = ON.warp_image_by_attention(imgs_h[b], full, H, H, "identity", remap_backend="cv2")
        diff = np.abs(out_h[b].astype(np.int16) - ref.astype(np.int16))
        worst = max(worst, int(diff.max()))
        n_off += int((diff != 0).sum())
    assert worst <= 1, f"worst image differs from the oracle by {worst} LSB"
    assert n_off <= 1e-4 * out_h.size, f"{n_off} of {out_h.size} bytes differ from the oracle"


def test_full_size_properties_c2():
    """BASELINE configs[1] at full size (256 x 336^2, [256,32,32,576] bf16): properties that do
    not need the CPU oracle on every image + oracle spot checks."""
    need_gpu()
    from attwarp_b200 import ops
    B, L, Hh, g, H = 256, 32, 32, 24, 336
    gen = torch.Generator(device="cuda").manual_seed(1)
    attn = torch.softmax(torch.randn(B, L, Hh, g * g, device="cuda", generator=gen) * 2, -1).to(torch.bfloat16)
    imgs = torch.randint(0, 256, (B, H, H, 3), device="cuda", dtype=torch.uint8, generator=gen)
    out, tok, mx, my = ops.warp_from_attention_tokens(attn, imgs, (g, g), transform="identity",
                                                      return_aux=True)
    # token maps are probability maps: every row of stage 1 sums to 1 -> so does the mean
    assert torch.allclose(tok.reshape(B, -1).sum(1), torch.ones(B, device="cuda"), atol=1e-4)
    # maps are monotone non-decreasing and inside [0, W]
    assert (mx[:, 1:] >= mx[:, :-1]).all() and (my[:, 1:] >= my[:, :-1]).all()
    assert mx.min() >= 0 and mx.max() <= H and my.min() >= 0 and my.max() <= H
    # determinism: a second run is bitwise identical
    out2 = ops.warp_from_attention_tokens(attn, imgs, (g, g), transform="identity")
    assert torch.equal(out, out2)
    # uniform attention + same output size is the identity warp
    uni = torch.full((B, 2, 2, g * g), 1.0 / (g * g), device="cuda", dtype=torch.bfloat16)
    ident = ops.warp_from_attention_tokens(uni, imgs, (g, g), transform="identity")
    assert torch.equal(ident, imgs)
    # EVERY image of the batch against the oracle (float64 stages 2-4 on the GPU's fp32 token map, the real
    # cv2.remap for stage 5): +-1 LSB, BASELINE.md section 4
    _check_all_images(out, imgs, tok, H)


def test_full_size_properties_c3():
    """BASELINE configs[2]: 64 x 1344^2 with a 48x48 token map."""
    need_gpu()
    from attwarp_b200 import ops
    B, g, H = 64, 48, 1344
    gen = torch.Generator(device="cuda").manual_seed(2)
    tok = torch.rand(B, g, g, device="cuda", generator=gen) ** 3
    tok = tok / tok.sum(dim=(1, 2), keepdim=True)
    imgs = torch.randint(0, 256, (B, H, H, 3), device="cuda", dtype=torch.uint8, generator=gen)
    mx, my = ops.maps_from_tokens(tok, (H, H), transform="identity")
    out = ops.remap_bilinear(imgs, mx, my, "hwc")
    assert (mx[:, 1:] >= mx[:, :-1]).all() and (my[:, 1:] >= my[:, :-1]).all()
    uni = torch.full_like(tok, 1.0 / (g * g))
    ux, uy = ops.maps_from_tokens(uni, (H, H), transform="identity")
    assert torch.equal(ops.remap_bilinear(imgs, ux, uy, "hwc"), imgs)
    _check_all_images(out, imgs, tok, H)


def test_stream_ring_matches_serial():
    """Consecutive batches spread over four streams (graph replays, concurrent kernels of different
    batches on the same SMs) give bit-identical results to the same batches run in turn."""
    need_gpu()
    from attwarp_b200 import ops
    from attwarp_b200.batched import StreamRing
    B, L, Hh, g, H, n = 32, 8, 8, 24, 336, 8
    gen = torch.Generator(device="cuda").manual_seed(7)
    sets = []
    for _ in range(n):
        attn = torch.softmax(torch.randn(B, L, Hh, g * g, device="cuda", generator=gen) * 2, -1).to(torch.bfloat16)
        img = torch.randint(0, 256, (B, H, H, 3), device="cuda", dtype=torch.uint8, generator=gen)
        aux = (torch.empty(B, g * g, device="cuda"), torch.empty(B, H, device="cuda"), torch.empty(B, H, device="cuda"))
        sets.append(dict(attn=attn, img=img, out=torch.zeros_like(img), aux=aux))

    def enqueue(s):
        ops.warp_from_attention_tokens(s["attn"], s["img"], (g, g), None, "hwc", out=s["out"], aux=s["aux"])

    serial = []
    for s in sets:
        enqueue(s)
        serial.append((s["out"].clone(), s["aux"][0].clone(), s["aux"][1].clone()))
        s["out"].zero_()
    torch.cuda.synchronize()
    graphs = [ops.GraphedCall(lambda s=s: enqueue(s)) for s in sets]
    for s in sets:
        s["out"].zero_()
    ring = StreamRing(4)
    ring.fork()
    for rep in range(3):
        for gph in graphs:
            ring.submit(gph.replay)
    ring.join()
    torch.cuda.synchronize()
    for s, (o, t, mx) in zip(sets, serial):
        assert torch.equal(s["out"], o)
        assert torch.equal(s["aux"][0], t)
        assert torch.equal(s["aux"][1], mx)


def _host_batch(B, L, Hh, g, H, seed):
    attn = _attention(B, L, Hh, g * g, torch.bfloat16, seed=seed)
    rng = np.random.default_rng(seed)
    imgs = torch.from_numpy(rng.integers(0, 256, (B, H, H, 3), dtype=np.uint8))
    return attn.pin_memory(), imgs.pin_memory()


def _check_host_batch(out_host, tok_host, attn, imgs, g, H, Ho):
    """Stage-wise, like test_fused_batch_vs_oracle: stage 1 against the oracle (1e-5 relative), stages 2-5 of the
    oracle (float64, real cv2.remap) fed with the pipeline's own fp32 token map -> every image +-1 LSB."""
    tok_ref = OA.aggregate_attention(attn.float().numpy())
    tok = tok_host.numpy()
    assert rel_err(tok, tok_ref, floor=1e-12) <= 1e-5
    worst = 0
    for b in range(attn.shape[0]):
        full = ON.upsample_tokens_nearest(tok[b].reshape(g, g), H, H)
        ref = ON.warp_image_by_attention(imgs[b].numpy(), full, Ho, Ho, "identity", remap_backend="cv2")
        diff = np.abs(out_host[b].numpy().astype(np.int16) - ref.astype(np.int16))
        worst = max(worst, int(diff.max()))
        assert (diff != 0).mean() <= 1e-3, f"image {b}: {(diff != 0).mean():.2e} of the bytes differ"
    assert worst <= 1, f"worst image differs from the oracle by {worst} LSB"


@pytest.mark.parametrize("B,chunk,Ho", [(37, 16, 336), (16, 16, 336), (5, 8, 500), (70, 32, 336)])
def test_host_batch_pipeline_vs_oracle(B, chunk, Ho):
    """The e2e API bench.py reports (pinned host -> H2D -> stages 1-5 -> D2H, chunks alternating over two
    streams and two device slots): B not a multiple of the chunk (partial last chunk), B < chunk, more chunks
    than slots (slot reuse), an output size different from the input -- every image within +-1 LSB of the
    oracle; then a SECOND run() of the same pipeline object on different data (slot and workspace reuse
    across calls) and a third on the first batch again, which must reproduce the first result bit for bit."""
    need_gpu()
    from attwarp_b200.batched import HostBatchPipeline
    L, Hh, g, H = 4, 8, 24, 336
    pipe = HostBatchPipeline(chunk, L, Hh, (g, g), (H, H, 3), out_hw=(Ho, Ho))
    attn1, imgs1 = _host_batch(B, L, Hh, g, H, seed=100 + B)
    attn2, imgs2 = _host_batch(B, L, Hh, g, H, seed=200 + B)
    out1 = torch.zeros(B, Ho, Ho, 3, dtype=torch.uint8).pin_memory()
    out2 = torch.zeros_like(out1).pin_memory()
    out3 = torch.zeros_like(out1).pin_memory()
    tok1 = torch.zeros(B, g * g).pin_memory()
    tok2 = torch.zeros(B, g * g).pin_memory()
    pipe.run(attn1, imgs1, out1, tok1)
    pipe.run(attn2, imgs2, out2, tok2)        # enqueued behind the first run, same slots
    pipe.run(attn1, imgs1, out3)
    pipe.sync()
    assert pipe.kernel_launches == 3 * 3 * ((B + chunk - 1) // chunk)
    _check_host_batch(out1, tok1, attn1, imgs1, g, H, Ho)
    _check_host_batch(out2, tok2, attn2, imgs2, g, H, Ho)
    assert torch.equal(out1, out3)
    assert not torch.equal(out1, out2)


def test_host_batch_pipeline_matches_device_path():
    """Same batch through the host pipeline and through the device-resident fused call: bit-identical."""
    need_gpu()
    from attwarp_b200 import ops
    from attwarp_b200.batched import HostBatchPipeline
    B, L, Hh, g, H, chunk = 50, 8, 8, 24, 336, 16
    attn, imgs = _host_batch(B, L, Hh, g, H, seed=9)
    out = torch.zeros(B, H, H, 3, dtype=torch.uint8).pin_memory()
    pipe = HostBatchPipeline(chunk, L, Hh, (g, g), (H, H, 3))
    pipe.run(attn, imgs, out)
    pipe.sync()
    ref = ops.warp_from_attention_tokens(attn.cuda(), imgs.cuda(), (g, g), transform="identity")
    assert torch.equal(out.cuda(), ref)


def test_fused_path_on_second_device():
    """kernel attributes (dynamic shared memory opt-in) are per device -- run the fused c2-shaped path
    on cuda:0 and then on cuda:1 from the same thread."""
    need_gpu()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from attwarp_b200 import ops
    B, L, Hh, g, H = 4, 32, 32, 24, 336
    attn = _attention(B, L, Hh, g * g, torch.bfloat16, seed=3)
    imgs = torch.from_numpy(np.random.default_rng(3).integers(0, 256, (B, H, H, 3), dtype=np.uint8))
    outs = []
    for d in (0, 1, 0):
        dv = torch.device("cuda", d)
        outs.append(ops.warp_from_attention_tokens(attn.to(dv), imgs.to(dv), (g, g), transform="identity").cpu())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


def test_sm_share_changes_grids_not_results():
    """attwarp_set_sm_share(2) (half an SM per launch of stage 1 / stage 5, for co-resident batches on several
    streams) only changes launch geometry: stage 5 gives the same bytes for the same maps, stage 1 the same token
    maps up to the float32 order of its split sums; StreamRing(sm_share=2) restores the previous mode."""
    need_gpu()
    from attwarp_b200 import _lib, ops
    from attwarp_b200.batched import StreamRing
    B, L, Hh, g, H = 40, 32, 32, 24, 336
    attn = _attention(B, L, Hh, g * g, torch.bfloat16, seed=5).cuda()
    imgs = torch.from_numpy(np.random.default_rng(5).integers(0, 256, (B, H, H, 3), dtype=np.uint8)).cuda()
    tok_ref = ops.aggregate_attention(attn)
    mx, my = ops.maps_from_tokens(tok_ref.view(B, g, g), (H, H))
    ref = ops.remap_bilinear(imgs, mx, my, "hwc")
    lib = _lib.load()
    prev = lib.attwarp_set_sm_share(2)
    try:
        tok = ops.aggregate_attention(attn)
        out = ops.remap_bilinear(imgs, mx, my, "hwc")
    finally:
        assert lib.attwarp_set_sm_share(prev) == 2
    assert torch.equal(out, ref)
    assert rel_err(tok.cpu().numpy(), tok_ref.cpu().numpy(), floor=1e-12) <= 1e-6
    ring = StreamRing(3, sm_share=2)
    ring.fork()
    outs = [ring.submit(lambda: ops.remap_bilinear(imgs, mx, my, "hwc")) for _ in range(3)]
    fused = ring.submit(lambda: ops.warp_from_attention_tokens(attn, imgs, (g, g)))
    ring.join()
    torch.cuda.synchronize()
    assert lib.attwarp_set_sm_share(1) == 1
    for o in outs:
        assert torch.equal(o, ref)
    d = (fused.to(torch.int16) - ref.to(torch.int16)).abs()
    assert int(d.max()) <= 8 and float((d != 0).float().mean()) <= 5e-3     # a few rint(32 x) flips from the split order
