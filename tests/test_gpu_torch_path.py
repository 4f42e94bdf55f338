"""GPU parity, torch path (safe_softmax .. warp_from_cdf_torch) and stage 1, against the golden
vectors recorded from the real reference.  Tolerances: BASELINE.md section 4."""

import numpy as np
import pytest
import torch

from gpu_util import dev, need_gpu, rel_err
from oracle import aggregate as OA
from oracle import torch_path as OT

pytestmark = pytest.mark.gpu


def test_softmax_mix(golden_torch):
    need_gpu()
    from attwarp_b200 import model
    g = golden_torch
    p = model.safe_softmax(dev(g["softmax/logits"]))
    assert rel_err(p.cpu().numpy(), g["softmax/p"]) <= 1e-5
    m = model.mix_with_uniform(dev(g["softmax/p"]), float(g["mix/alpha"]))
    assert rel_err(m.cpu().numpy(), g["mix/p"]) <= 1e-5
    same = dev(g["softmax/p"])
    assert model.mix_with_uniform(same, 0.0) is same
    # other dims / shapes
    z = torch.randn(4, 7, 5, device="cuda")
    ref = OT.safe_softmax(z.permute(0, 2, 1).reshape(-1, 7).cpu().numpy()).reshape(4, 5, 7)
    got = model.safe_softmax(z, dim=1).permute(0, 2, 1).cpu().numpy()
    assert rel_err(got, ref) <= 1e-5


@pytest.mark.parametrize("L", [336, 512, 100])
def test_upsample_and_cdf(golden_torch, L):
    need_gpu()
    from attwarp_b200 import checkpoint_utils as cu
    g = golden_torch
    up = cu.upsample_pdf_right_inverse(dev(g["softmax/p"]), L).cpu().numpy()
    # absolute floor = the reference's own fp32 LAPACK noise (see test_oracle_vs_golden.py)
    assert np.abs(up - g[f"upsample/{L}"]).max() <= 1e-6
    assert np.abs(up - g[f"upsample64/{L}"]).max() <= 2e-7      # closer to the fp64 truth
    F = cu.cdf_from_density(dev(np.maximum(g[f"upsample/{L}"], 0))).cpu().numpy()
    assert rel_err(F, g[f"cdf_from_upsample/{L}"]) <= 1e-5
    assert (np.diff(F, axis=1) >= 0).all() and (F[:, -1] == 1).all()


def test_upsample_shapes_and_errors(golden_torch):
    need_gpu()
    from attwarp_b200 import checkpoint_utils as cu
    g = golden_torch
    p = g["softmax/p"]
    assert np.abs(cu.upsample_pdf_right_inverse(dev(p[0]), 48).cpu().numpy() - g["upsample/1d"]).max() <= 1e-6
    got = cu.upsample_pdf_right_inverse(dev(p.reshape(4, 4, 24)), 48).cpu().numpy()
    assert got.shape == (4, 4, 48) and np.abs(got - g["upsample/3d"]).max() <= 1e-6
    with pytest.raises(ValueError):
        cu.upsample_pdf_right_inverse(torch.zeros(1, 1, 1, 24, device="cuda"), 48)


def test_cdf_marginals_pool(golden_torch):
    need_gpu()
    from attwarp_b200 import checkpoint_utils as cu
    g = golden_torch
    F = cu.cdf_from_density(dev(g["cdf/p"])).cpu().numpy()
    assert rel_err(F, g["cdf/F"]) <= 1e-5
    mx, my = cu.gt_marginals(dev(g["gt/A"]))
    assert rel_err(mx.cpu().numpy(), g["gt/mx"]) <= 1e-5
    assert rel_err(my.cpu().numpy(), g["gt/my"]) <= 1e-5
    pooled = cu.adaptive_avg_pool2d_24(dev(g["pool/A"])).cpu().numpy()
    assert rel_err(pooled, g["pool/out"]) <= 1e-5
    big = np.random.default_rng(0).random((2, 1, 512, 512)).astype(np.float32)
    ref = torch.nn.functional.adaptive_avg_pool2d(torch.from_numpy(big), (24, 24)).numpy()
    assert rel_err(cu.adaptive_avg_pool2d_24(dev(big)).cpu().numpy(), ref) <= 1e-5
    rx, ry = OT.gt_marginals(big)
    mx, my = cu.gt_marginals(dev(big))
    assert rel_err(mx.cpu().numpy(), rx) <= 1e-5 and rel_err(my.cpu().numpy(), ry) <= 1e-5
    # odd shapes: scalar path (W % 4 != 0), windows of one pixel, up-sampling windows (H < gh)
    for (H, W) in [(97, 53), (24, 24), (10, 36), (336, 500)]:
        a = np.random.default_rng(H + W).random((3, 1, H, W)).astype(np.float32)
        ref = torch.nn.functional.adaptive_avg_pool2d(torch.from_numpy(a), (24, 24)).numpy()
        assert rel_err(cu.adaptive_avg_pool2d_24(dev(a)).cpu().numpy(), ref) <= 1e-5, (H, W)


def test_resample_cdf_and_strictly_increasing(golden_torch):
    """checkpoint_utils.py:17-28, 53-62 (plot-only callers in the reference) against the reference's
    outputs, plus NaN / Inf / flat / decreasing rows against the oracle."""
    need_gpu()
    from attwarp_b200 import checkpoint_utils as cu
    g = golden_torch
    assert rel_err(cu._make_strictly_increasing(dev(g["resample/F"])).cpu().numpy(), g["strict/out"]) <= 1e-5
    assert rel_err(cu.resample_cdf(dev(g["resample/F"]), 200).cpu().numpy(), g["resample/out"]) <= 1e-5
    rng = np.random.default_rng(8)
    F = np.sort(rng.random((6, 700)).astype(np.float32), axis=1)
    F[0, 5] = np.nan; F[1, 9] = np.inf; F[1, 3] = -np.inf; F[2, :] = 0.25; F[3] = F[3, ::-1]
    got = cu._make_strictly_increasing(dev(F)).cpu().numpy()
    assert rel_err(got, OT.make_strictly_increasing(F)) <= 1e-5
    assert np.all(np.diff(got, axis=1) > 0) and np.all(got[:, -1] == 1.0)
    for L in (64, 700, 1500):
        assert rel_err(cu.resample_cdf(dev(F), L).cpu().numpy(), OT.resample_cdf(F, L)) <= 1e-5
    out = cu.resample_cdf(torch.from_numpy(F[:2]), 100)
    assert out.device.type == "cpu" and out.shape == (2, 100)


@pytest.mark.parametrize("name", ["u8_same", "f32_out", "u8_odd", "f32_c4", "u8_c4_sharp"])
def test_warp_from_cdf_torch(golden_torch, name):
    need_gpu()
    from attwarp_b200 import checkpoint_utils as cu, ops
    g = golden_torch
    osz = tuple(int(v) for v in g[f"warp/{name}/out_size"])
    osz = None if osz[0] < 0 else osz
    img, Fx, Fy = g[f"warp/{name}/img"], g[f"warp/{name}/Fx"], g[f"warp/{name}/Fy"]
    ref = g[f"warp/{name}/out"]
    # stage 4 fed with the reference's CDFs: maps within 1e-4 px of the oracle's
    B, C, H, W = img.shape
    Ho, Wo = (H, W) if osz is None else osz
    mx, my = ops.maps_from_cdf(dev(Fx), dev(Fy), (Ho, Wo))
    rx, ry = OT.maps_from_cdf(Fx, Fy, (Ho, Wo))
    assert np.abs(mx.cpu().numpy() - rx).max() <= 1e-4 and np.abs(my.cpu().numpy() - ry).max() <= 1e-4
    # CPU tensors in -> CPU tensor out, like the reference (device round trip inside)
    out = cu.warp_from_cdf_torch(torch.from_numpy(img), torch.from_numpy(Fx), torch.from_numpy(Fy), osz)
    assert out.device.type == "cpu" and out.dtype == torch.from_numpy(img).dtype
    out = out.numpy()
    assert out.shape == ref.shape
    if ref.dtype == np.uint8:
        assert np.array_equal(out, ref)
    else:
        assert np.abs(out - ref).max() <= 1e-6
    out_gpu = cu.warp_from_cdf_torch(dev(img), dev(Fx), dev(Fy), osz)
    assert out_gpu.is_cuda and np.array_equal(out_gpu.cpu().numpy(), out)


def test_warp_from_cdf_errors():
    need_gpu()
    from attwarp_b200 import checkpoint_utils as cu
    img = torch.zeros(1, 3, 8, 9, dtype=torch.uint8, device="cuda")
    with pytest.raises(ValueError):
        cu.warp_from_cdf_torch(img, torch.zeros(1, 8, device="cuda"), torch.zeros(1, 8, device="cuda"))
    with pytest.raises(ValueError):
        cu.warp_from_cdf_torch(img, torch.zeros(1, 9, device="cuda"), torch.zeros(1, 9, device="cuda"))
    with pytest.raises(AssertionError):
        cu.warp_from_cdf_torch(img[0], torch.zeros(1, 9, device="cuda"), torch.zeros(1, 8, device="cuda"))
    # single-channel works here (the reference crashes, documented superset)
    one = cu.warp_from_cdf_torch(img[:, :1], torch.linspace(0.1, 1, 9, device="cuda")[None],
                                 torch.linspace(0.1, 1, 8, device="cuda")[None])
    assert one.shape == (1, 1, 8, 9)


# ------------------------------------------------------------------------------ stage 1
@pytest.mark.parametrize("dtype", ["f32", "bf16", "f16"])
def test_aggregate_golden(golden_aggregate, dtype):
    need_gpu()
    from attwarp_b200 import ops
    g = golden_aggregate
    T = int(g["T"])
    starts = g["starts"]
    a = torch.from_numpy(g["attn_f32"]).cuda()
    if dtype == "bf16":
        a = torch.from_numpy(g["attn_bf16_bits"]).cuda().view(torch.bfloat16)
        ref = g["batch_logger/bf16"]
    elif dtype == "f16":
        a = a.half()
        ref = OA.aggregate_attention(a.float().cpu().numpy(), starts, T)
    else:
        ref = g["batch_logger/f32"]
    out = ops.aggregate_attention(a, tok_start=torch.from_numpy(starts), num_tokens=T)
    assert rel_err(out.cpu().numpy(), ref) <= 1e-5
    # contiguous pre-sliced layout -> vectorised kernel
    sl = torch.stack([a[b, :, :, int(starts[b]):int(starts[b]) + T] for b in range(a.shape[0])]).contiguous()
    out2 = ops.aggregate_attention(sl)
    assert rel_err(out2.cpu().numpy(), ref) <= 1e-5


@pytest.mark.parametrize("B,L,Hh,T,dt", [(5, 32, 32, 576, torch.bfloat16), (2, 3, 5, 576, torch.float32),
                                         (3, 4, 8, 2304, torch.bfloat16), (2, 2, 4, 1000, torch.float16),
                                         (1, 1, 32, 576, torch.float16), (4, 7, 3, 24, torch.float32)])
def test_aggregate_shapes_vs_oracle(B, L, Hh, T, dt):
    need_gpu()
    from attwarp_b200 import ops
    gen = torch.Generator().manual_seed(B * 100 + T)
    a = torch.softmax(torch.randn(B, L, Hh, T + 37, generator=gen), dim=-1)[..., :T].contiguous().to(dt)
    ref = OA.aggregate_attention(a.float().numpy())
    out = ops.aggregate_attention(a.cuda())
    assert rel_err(out.cpu().numpy(), ref) <= 1e-5


def test_hook_loggers(golden_aggregate):
    need_gpu()
    from attwarp_b200.attention_extraction import BatchMaskHookLogger, MaskHookLogger
    g = golden_aggregate
    a = torch.from_numpy(g["attn_f32"]).cuda()
    B, L, Hh, K = a.shape
    T = int(g["T"])
    starts = [int(s) for s in g["starts"]]
    bl = BatchMaskHookLogger(None, "cuda")
    bl.set_batch_image_token_ranges(starts, [s + T for s in starts])
    for l in range(L):
        q = torch.zeros(B, Hh, 2, K, device="cuda")
        q[:, :, -1, :] = a[:, l]
        bl._process_attention(q)
    maps = bl.finalize_batch()
    assert len(maps) == B and maps[0].shape == (24, 24)
    got = torch.stack([m.reshape(-1) for m in maps]).cpu().numpy()
    assert rel_err(got, g["batch_logger/f32"]) <= 1e-5
    ml = MaskHookLogger(None, "cuda")
    ml.set_image_token_range(starts[1], starts[1] + T)
    for l in range(L):
        q = torch.zeros(1, Hh, 2, K, device="cuda")
        q[:, :, -1, :] = a[1:2, l]
        ml._process_attention(q)
    assert rel_err(ml.finalize().cpu().numpy(), g["single_logger/f32"]) <= 1e-5
    ml.reinit()
    q = torch.zeros(1, Hh, 3, K, device="cuda")
    q[:, :, -1, :] = a[0:1, 0]
    ml._process_attention(q)                      # default token range 1..577
    assert rel_err(ml.finalize().cpu().numpy(), g["single_logger/default_range"]) <= 1e-5
    ml.reinit()
    assert rel_err(ml.finalize().cpu().numpy(), g["single_logger/empty"]) <= 1e-6


def test_hook_logger_factories_on_a_stub_model():
    """hook_logger / batch_hook_logger (llava.py:156-187, 451-462) on a stand-in for the decoder: the
    registered forward hook must pick the attention weights out of the layer's output tuple and reduce
    them on the device exactly like _process_attention fed by hand."""
    need_gpu()
    import types
    from attwarp_b200.attention_extraction import (BatchMaskHookLogger, MaskHookLogger, batch_hook_logger,
                                                  hook_logger)

    class Attn(torch.nn.Module):
        def forward(self, attn, output_attentions=False):
            return (attn.sum(), attn if output_attentions else None, None)

    def make_model(n_layers=3):
        layers = torch.nn.ModuleList([torch.nn.Module() for _ in range(n_layers)])
        for l in layers:
            l.self_attn = Attn()
        m = torch.nn.Module()
        m.model = torch.nn.Module()
        m.model.layers = layers
        m.config = types.SimpleNamespace(output_attentions=False)
        return m

    gen = torch.Generator().manual_seed(11)
    B, Hh, q, kv = 3, 4, 5, 640
    attn = torch.softmax(torch.randn(B, Hh, q, kv, generator=gen), -1).cuda()
    starts, ends = [1, 7, 30], [577, 583, 606]

    model = make_model()
    bl = batch_hook_logger(model, "cuda", layer_index=1)
    assert isinstance(bl, BatchMaskHookLogger) and model.batch_hooklogger is bl
    assert model.config.output_attentions is False
    bl.set_batch_image_token_ranges(starts, ends)
    for _ in range(2):
        model.model.layers[1].self_attn(attn)             # the patch forces output_attentions=True
    model.model.layers[0].self_attn(attn)                  # other layers are not hooked
    ref = BatchMaskHookLogger(None, "cuda")
    ref.set_batch_image_token_ranges(starts, ends)
    for _ in range(2):
        ref._process_attention(attn)
    for a, b in zip(bl.finalize_batch(), ref.finalize_batch()):
        assert torch.equal(a, b)
    bl.remove_hook_and_unpatch()

    model = make_model()
    sl = hook_logger(model, "cuda", layer_index=2)
    assert isinstance(sl, MaskHookLogger) and model.hooklogger is sl and model.config.output_attentions is True
    sl.set_image_token_range(7, 583)
    model.model.layers[2].self_attn(attn[1:2], output_attentions=True)
    r2 = MaskHookLogger(None, "cuda")
    r2.set_image_token_range(7, 583)
    r2._process_attention(attn[1:2])
    assert torch.equal(sl.finalize(), r2.finalize())
    sl.remove_hook()


def test_pool_attention_prologue():
    """trainer.py:172-197: clamp_min(0), per-sample sqrt mask, adaptive_avg_pool2d -- fused, against the
    reference's own torch expressions."""
    need_gpu()
    from attwarp_b200 import checkpoint_utils as cu
    gen = torch.Generator().manual_seed(21)
    for (H, W) in [(512, 512), (97, 54)]:
        A = torch.randn(5, 1, H, W, generator=gen) * 0.5 + 0.3          # some negatives
        tf = ["sqrt", "iden", "none", "sqrt", "iden"]
        m = torch.tensor([1.0 if t == "sqrt" else 0.0 for t in tf]).view(5, 1, 1, 1)
        pos = A.clamp_min(0.0)
        ref = torch.nn.functional.adaptive_avg_pool2d(pos.sqrt() * m + pos * (1.0 - m), (24, 24)).numpy()
        got = cu.pool_attention(A.cuda(), tf).cpu().numpy()
        assert rel_err(got, ref, floor=1e-6) <= 1e-5, (H, W)
        plain = cu.pool_attention(A.cuda(), None).cpu().numpy()
        assert rel_err(plain, torch.nn.functional.adaptive_avg_pool2d(A, (24, 24)).numpy(), floor=1e-6) <= 1e-4


def test_hook_logger_prepared_call_with_growing_kv(monkeypatch):
    """After the first step of a layout the hook reducer keeps a prepared C call (attention_extraction.py:
    _RunningAttention._prepare).  Under generate() the tensor changes from step to step -- q = prompt length, then 1;
    kv grows with the KV cache; the views are not contiguous -- so the prepared call reads strides and the last query
    row per step: same result, bit for bit, as sending every step through ops.aggregate_attention, and equal to the
    definition (llava.py:385-411) computed with torch."""
    need_gpu()
    from attwarp_b200 import attention_extraction as AE
    gen = torch.Generator().manual_seed(5)
    B, Hh, T = 3, 4, 576
    starts, ends = [1, 7, 30], [577, 583, 606]
    big = torch.softmax(torch.randn(B, Hh, 5, 700, generator=gen), -1).cuda().half()
    steps = [big[:, :, :, :640]] + [big[:, :, 4 - (i % 3):5 - (i % 3), :641 + 3 * i] for i in range(9)]
    results = []
    for fast in (True, False):
        monkeypatch.setattr(AE, "_HOOK_FAST", fast)
        lg = AE.BatchMaskHookLogger(None, "cuda")
        lg.set_batch_image_token_ranges(starts, ends)
        for s in steps:
            lg._process_attention(s)
        assert (lg._acc._fast is not None) == fast
        results.append(torch.stack([m.reshape(-1) for m in lg.finalize_batch()]))
    assert torch.equal(results[0], results[1])
    ref = torch.zeros(B, T, device="cuda")
    for s in steps:
        row = s[:, :, -1, :].float()
        for b in range(B):
            sl = row[b, :, starts[b]:ends[b]]
            ref[b] += (sl / (sl.sum(-1, keepdim=True) + 1e-12)).mean(0)
    ref /= len(steps)
    assert rel_err(results[0].cpu().numpy(), ref.cpu().numpy()) <= 1e-5
    # a change of batch size or of the ranges leaves the prepared call
    lg = AE.MaskHookLogger(None, "cuda")
    lg.set_image_token_range(1, 577)
    lg._process_attention(big[:1, :, :, :640])
    lg._process_attention(big[:1, :, -1:, :650])
    with pytest.raises(ValueError):
        lg._process_attention(big[:2, :, -1:, :650])          # running sum is [1, T]
