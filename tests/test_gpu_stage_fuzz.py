"""GPU parity, stages 1-4 and the driver-flow mask: seeded random shapes, dtypes, strides and parameters against the
oracle (and Pillow for the LANCZOS resize).  Tolerances as in BASELINE.md section 4.  ATTWARP_FUZZ_CASES raises the
number of cases (default 48)."""

import os

import numpy as np
import pytest
import torch

from gpu_util import dev, need_gpu, rel_err
from oracle import aggregate as OA
from oracle import numpy_path as ON
from oracle import torch_path as OT

pytestmark = pytest.mark.gpu

N_CASES = int(os.environ.get("ATTWARP_FUZZ_CASES", "48"))
TRANSFORMS = ["identity", "sqrt", "square", "exp", "log"]


@pytest.mark.parametrize("case", range(max(8, N_CASES // 2)))
def test_aggregate_random(case):
    """Stage 1: random [B, L, Hh, K] shapes, dtypes, token offsets, dense and strided (live-hook) layouts."""
    need_gpu()
    from attwarp_b200 import ops
    rng = np.random.default_rng(7000 + case)
    gen = torch.Generator().manual_seed(7000 + case)
    B, L, Hh = int(rng.integers(1, 7)), int(rng.integers(1, 9)), int(rng.integers(1, 33))
    T = int(rng.choice([1, 7, 24, 100, 576, 577, 1000, 2304]))
    extra = int(rng.integers(0, 90))
    dt = [torch.float32, torch.bfloat16, torch.float16][case % 3]
    a = torch.softmax(torch.randn(B, L, Hh, T + extra, generator=gen) * 2, dim=-1).to(dt)
    starts = rng.integers(0, extra + 1, size=B).astype(np.int32)
    ref = OA.aggregate_attention(a.float().numpy(), starts, T)
    out = ops.aggregate_attention(a.cuda(), tok_start=torch.from_numpy(starts), num_tokens=T)
    assert rel_err(out.cpu().numpy(), ref) <= 1e-5, (case, B, L, Hh, T, extra, dt)
    if L == 1:
        # the live-hook layout: [B, Hh, q, kv] addressed at the last query row through strides (no copy)
        q = int(rng.integers(1, 5))
        full = torch.zeros(B, Hh, q, T + extra, dtype=dt)
        full[:, :, -1, :] = a[:, 0]
        view = full.cuda()[:, :, -1, :].unsqueeze(1)                  # [B, 1, Hh, K], strided
        out2 = ops.aggregate_attention(view, tok_start=torch.from_numpy(starts), num_tokens=T)
        assert rel_err(out2.cpu().numpy(), ref) <= 1e-5, (case, "strided")


@pytest.mark.parametrize("case", range(N_CASES))
def test_maps_from_tokens_random(case):
    """Stages 2-4 from a token grid: random grids, image and output sizes, transforms, inverse on the marginals."""
    need_gpu()
    from attwarp_b200 import ops
    rng = np.random.default_rng(7300 + case)
    gh, gw = int(rng.integers(1, 50)), int(rng.integers(1, 50))
    H, W = int(rng.integers(max(gh, 2), 900)), int(rng.integers(max(gw, 2), 1500))
    Ho, Wo = int(rng.integers(1, 900)), int(rng.integers(1, 1500))
    tr = TRANSFORMS[case % len(TRANSFORMS)]
    inv = bool(rng.integers(0, 2))
    es, ed = float(rng.uniform(0.5, 3.0)), float(rng.uniform(0.5, 4.0))
    B = int(rng.integers(1, 4))
    tok = (rng.random((B, gh, gw)) ** int(rng.integers(1, 5))).astype(np.float32)
    if case % 11 == 0:
        tok[0] = 0.0                                                   # near-zero attention: the uniform fallback
    mx, my = ops.maps_from_tokens(dev(tok), (H, W), (Ho, Wo), tr, es, ed, inv)
    for b in range(B):
        full = ON.upsample_tokens_nearest(tok[b], H, W)
        rx, ry, _, _ = ON.inverse_maps(full, Wo, Ho, tr, es, ed, inv)
        ex = np.abs(mx[b].cpu().numpy() - rx.astype(np.float32)).max()
        ey = np.abs(my[b].cpu().numpy() - ry.astype(np.float32)).max()
        assert ex <= 1e-4 * max(1.0, W / 1000) and ey <= 1e-4 * max(1.0, H / 1000), (case, (gh, gw), (H, W), (Ho, Wo), tr, inv, ex, ey)


@pytest.mark.parametrize("case", range(N_CASES))
def test_maps_from_attention_random(case):
    """Stages 2b-4 from a materialised map: uint8 / float32 / float64, every transform, odd sizes and alignments."""
    need_gpu()
    from attwarp_b200 import ops
    rng = np.random.default_rng(7600 + case)
    H, W = int(rng.integers(1, 700)), int(rng.integers(1, 1800))
    if case % 3 == 0:
        W = (W + 15) & ~15                                             # the uint8 / float32 fast paths need aligned rows
    Ho, Wo = int(rng.integers(1, 700)), int(rng.integers(1, 1800))
    tr = TRANSFORMS[(case // 3) % len(TRANSFORMS)]
    inv = bool(rng.integers(0, 2))
    B = int(rng.integers(1, 4))
    kind = case % 3
    if kind == 0:
        att = rng.integers(0, 256, (B, H, W), dtype=np.uint8)
    elif kind == 1:
        att = (rng.random((B, H, W)) ** 2).astype(np.float32)
    else:
        att = rng.random((B, H, W)) ** 2 - 0.05                        # float64 with some negatives (clamped)
    mx, my = ops.maps_from_attention(dev(att), (Ho, Wo), tr, 1.5, 2.0, inv)
    for b in range(B):
        with np.errstate(all="ignore"):
            rx, ry, _, _ = ON.inverse_maps(att[b], Wo, Ho, tr, 1.5, 2.0, inv)
        if not (np.isfinite(rx).all() and np.isfinite(ry).all()):
            continue        # the reference itself overflows here (exp of a marginal SUM with apply_inverse): nothing to compare
        ex = np.abs(mx[b].cpu().numpy() - rx.astype(np.float32)).max()
        ey = np.abs(my[b].cpu().numpy() - ry.astype(np.float32)).max()
        assert ex <= 1e-4 * max(1.0, W / 1000) and ey <= 1e-4 * max(1.0, H / 1000), (case, att.dtype, (H, W), (Ho, Wo), tr, inv, ex, ey)


@pytest.mark.parametrize("case", range(max(8, N_CASES // 2)))
def test_maps_from_cdf_random(case):
    """Stage 4 of the torch path: random CDFs with flat stretches (the tie-break of checkpoint_utils.py:181-184)."""
    need_gpu()
    from attwarp_b200 import ops
    rng = np.random.default_rng(7900 + case)
    B, H, W = int(rng.integers(1, 5)), int(rng.integers(2, 700)), int(rng.integers(2, 1500))
    Ho, Wo = int(rng.integers(1, 700)), int(rng.integers(1, 1500))

    def cdf(n):
        p = rng.random((B, n)) ** 3
        p[rng.random((B, n)) < 0.3] = 0.0                              # flat stretches
        p[:, 0] += 1e-3
        return OT.cdf_from_density(p.astype(np.float32))

    Fx, Fy = cdf(W), cdf(H)
    mx, my = ops.maps_from_cdf(dev(Fx), dev(Fy), (Ho, Wo))
    rx, ry = OT.maps_from_cdf(Fx, Fy, (Ho, Wo))
    assert np.abs(mx.cpu().numpy() - rx).max() <= 1e-4 * max(1.0, W / 1000), (case, (H, W), (Ho, Wo))
    assert np.abs(my.cpu().numpy() - ry).max() <= 1e-4 * max(1.0, H / 1000), (case, (H, W), (Ho, Wo))


@pytest.mark.parametrize("case", range(max(8, N_CASES // 2)))
def test_lanczos_random(case):
    """The Pillow-exact LANCZOS resize of mode-'L' images: random sizes, up- and down-scaling, bit-equal to PIL; and
    the fused resize + marginals path against the two-step path."""
    need_gpu()
    from PIL import Image
    from attwarp_b200 import ops
    rng = np.random.default_rng(8200 + case)
    h, w = int(rng.integers(1, 60)), int(rng.integers(1, 60))
    if case % 4 == 3:
        h, w = int(rng.integers(60, 100)), int(rng.integers(60, 160))  # down-scaling / general kernel (the source and
        # its horizontal pass must fit shared memory: the masks of the driver flow are token grids)
    Ho, Wo = int(rng.integers(1, 900)), int(rng.integers(1, 1400))
    m = rng.integers(0, 256, (2, h, w), dtype=np.uint8)
    got = ops.resize_lanczos_u8(dev(m), (Ho, Wo)).cpu().numpy()
    for b in range(2):
        ref = np.array(Image.fromarray(m[b], mode="L").resize((Wo, Ho), Image.LANCZOS))
        assert np.array_equal(got[b], ref), (case, (h, w), (Ho, Wo))
    if Ho > h and Wo > w and h <= 60 and w <= 60:
        tok = torch.from_numpy(rng.random((2, max(h, 2), max(w, 2))).astype(np.float32)).cuda()
        a = ops.maps_from_mota_tokens(tok, (Ho, Wo), (Ho, Wo), fused=True)
        b_ = ops.maps_from_mota_tokens(tok, (Ho, Wo), (Ho, Wo), fused=False)
        for x, y in zip(a, b_):
            ulp = np.abs(x.cpu().numpy().view(np.int32).astype(np.int64) - y.cpu().numpy().view(np.int32).astype(np.int64))
            assert ulp.max() <= 1, (case, "fused vs two-step", int(ulp.max()))


@pytest.mark.parametrize("case", range(max(8, N_CASES // 2)))
def test_torch_helpers_random(case):
    """The torch-path helpers at random sizes: safe_softmax (with non-finite logits), mix_with_uniform and their fused
    form, cdf_from_density, the right-inverse upsample at lengths that are and are not multiples of the grid,
    gt_marginals and adaptive pooling at sizes that miss the vectorised paths."""
    need_gpu()
    from attwarp_b200 import checkpoint_utils as CU, model as M
    rng = np.random.default_rng(8500 + case)
    B, N = int(rng.integers(1, 200)), int(rng.integers(1, 80))
    z = (rng.standard_normal((B, N)) * 4).astype(np.float32)
    if case % 3 == 0:
        z[rng.random((B, N)) < 0.05] = np.nan
        z[rng.random((B, N)) < 0.05] = np.inf
        z[rng.random((B, N)) < 0.05] = -np.inf
    alpha = float(rng.choice([0.0, 0.03, 0.5]))
    with np.errstate(all="ignore"):
        p_ref = OT.safe_softmax(z)
        m_ref = OT.mix_with_uniform(p_ref, alpha)
    p = M.safe_softmax(dev(z))
    assert rel_err(p.cpu().numpy(), p_ref, 1e-7) <= 2e-5, (case, "safe_softmax", B, N)
    assert rel_err(M.mix_with_uniform(p, alpha).cpu().numpy(), OT.mix_with_uniform(p.cpu().numpy(), alpha), 1e-7) <= 1e-5
    assert rel_err(M.safe_softmax_mix(dev(z), alpha).cpu().numpy(), m_ref, 1e-7) <= 2e-5, (case, "softmax_mix")
    # CDFs
    d = (rng.random((B, N)) ** 2 - 0.1).astype(np.float32)
    assert np.abs(CU.cdf_from_density(dev(d)).cpu().numpy() - OT.cdf_from_density(d)).max() <= 1e-5, (case, "cdf")
    # right-inverse upsample
    L_out = int(rng.integers(2, 30))
    L_in = int(rng.choice([L_out * int(rng.integers(1, 30)), int(rng.integers(L_out, 700))]))
    y = rng.random((min(B, 16), L_out)).astype(np.float32)
    y /= y.sum(axis=1, keepdims=True)
    got = CU.upsample_pdf_right_inverse(dev(y), L_in).cpu().numpy()
    ref = OT.upsample_pdf_right_inverse(y, L_in)
    assert np.abs(got - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max()), (case, "upsample", L_out, L_in)
    # marginals and pooling of a full-resolution map
    H, W = int(rng.integers(1, 300)), int(rng.integers(1, 700))
    A = (rng.random((min(B, 3), 1, H, W)) - 0.2).astype(np.float32)
    px, py = CU.gt_marginals(dev(A))
    rx, ry = OT.gt_marginals(A)
    assert rel_err(px.cpu().numpy(), rx, 1e-9) <= 2e-5 and rel_err(py.cpu().numpy(), ry, 1e-9) <= 2e-5, (case, "gt_marginals", H, W)
    gh, gw = int(rng.integers(1, min(H, 30) + 1)), int(rng.integers(1, min(W, 30) + 1))
    got = CU.adaptive_avg_pool2d_24(dev(A), (gh, gw)).cpu().numpy()
    ref = OT.adaptive_avg_pool2d(A, (gh, gw))
    assert np.abs(got - ref).max() <= 1e-5, (case, "pool", (H, W), (gh, gw))
