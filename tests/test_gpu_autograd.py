"""The helpers the reference TRAINS through (mnfd/trainer.py:209-250: net -> safe_softmax -> mix_with_uniform ->
upsample_pdf_right_inverse -> clamp / normalise -> L1) must stay differentiable when the mirrors replace them:
gradients of the library's backward kernels against torch.autograd on the reference's own expressions
(restated below with their file:line), float32, <= 1e-5 relative to the largest gradient."""

import os

import numpy as np
import pytest
import torch

from gpu_util import need_gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "autograd.npz")


def ref_safe_softmax(logits, dim=1, eps=1e-6):          # model.py:8-14
    logits = torch.nan_to_num(logits, nan=0.0, posinf=0.0, neginf=0.0)
    logits = logits - logits.amax(dim=dim, keepdim=True)
    p = torch.softmax(logits, dim=dim)
    p = torch.nan_to_num(p, nan=0.0, posinf=0.0, neginf=0.0)
    return p / p.sum(dim=dim, keepdim=True).clamp_min(eps)


def ref_mix_with_uniform(p, alpha):                     # model.py:98-101
    if alpha <= 0:
        return p
    return (1 - alpha) * p + alpha / p.size(1)


def ref_upsample(y, L_in, eps=1e-8):                    # checkpoint_utils.py:64-131
    L_out = y.shape[-1]
    A = torch.zeros(L_out, L_in, dtype=torch.float64)
    for i in range(L_out):
        s, e = (i * L_in) // L_out, -((-(i + 1) * L_in) // L_out)
        A[i, s:e] = 1.0 / (e - s)
    G = A @ A.T + eps * torch.eye(L_out, dtype=torch.float64)
    z = torch.linalg.solve(G, y.double().T)
    return (A.T @ z).T.float()


def ref_cdf(p):                                         # checkpoint_utils.py:30-41
    p = p.clamp_min(0)
    p = torch.nan_to_num(p, nan=0.0, posinf=0.0, neginf=0.0)
    p = p / p.sum(dim=1, keepdim=True).clamp_min(1e-6)
    Fp = p.cumsum(dim=1)
    Fp = Fp.clone()
    Fp[:, -1] = 1.0
    return Fp


def _close(a, b, tol=1e-5):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max()) <= tol * max(float(b.abs().max()), 1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("alpha,L_in", [(0.0, 512), (0.1, 512), (0.3, 336)])
def test_training_chain_gradients(alpha, L_in):
    need_gpu()
    from attwarp_b200 import checkpoint_utils as CU, model as M
    g = torch.Generator().manual_seed(3)
    z = torch.randn(16, 24, generator=g) * 2
    gt = torch.softmax(torch.randn(16, L_in, generator=g), -1)

    def chain(ss, mix, up, zz, gtt):
        p = mix(ss(zz, dim=1, eps=1e-6), alpha)
        x = up(p, L_in).clamp_min(0)
        x = x / x.sum(dim=1, keepdim=True).clamp_min(1e-6)
        return torch.nn.functional.l1_loss(x, gtt), x

    zr = z.clone().requires_grad_(True)
    loss_r, x_r = chain(ref_safe_softmax, ref_mix_with_uniform, ref_upsample, zr, gt)
    loss_r.backward()
    zg = z.clone().cuda().requires_grad_(True)
    loss_g, x_g = chain(M.safe_softmax, M.mix_with_uniform, CU.upsample_pdf_right_inverse, zg, gt.cuda())
    loss_g.backward()
    assert _close(x_g, x_r) and abs(float(loss_g.detach()) - float(loss_r.detach())) <= 1e-6
    assert zg.grad is not None and _close(zg.grad, zr.grad, 2e-4)


@pytest.mark.gpu
def test_each_backward_kernel():
    need_gpu()
    from attwarp_b200 import checkpoint_utils as CU, model as M
    g = torch.Generator().manual_seed(5)
    # safe_softmax with a non-finite logit (zero slope there), 3-D input along dim 1
    z = torch.randn(4, 24, 3, generator=g)
    z[1, 5, 0] = float("inf")
    w = torch.randn(4, 24, 3, generator=g)
    zr = z.clone().requires_grad_(True)
    (ref_safe_softmax(zr, dim=1) * w).sum().backward()
    zg = z.clone().cuda().requires_grad_(True)
    (M.safe_softmax(zg, dim=1) * w.cuda()).sum().backward()
    assert _close(zg.grad, zr.grad, 1e-4)
    # mix_with_uniform
    p = torch.softmax(torch.randn(8, 24, generator=g), -1)
    w = torch.randn(8, 24, generator=g)
    pr = p.clone().requires_grad_(True)
    (ref_mix_with_uniform(pr, 0.25) * w).sum().backward()
    pg = p.clone().cuda().requires_grad_(True)
    (M.mix_with_uniform(pg, 0.25) * w.cuda()).sum().backward()
    assert _close(pg.grad, pr.grad)
    # upsample_pdf_right_inverse: divisible and non-divisible lengths, 1-D / 3-D inputs
    for shape, L_in in (((8, 24), 512), ((24,), 336), ((2, 3, 24), 100)):
        y = torch.rand(*shape, generator=g)
        w = torch.randn(*shape[:-1], L_in, generator=g)
        yr = y.clone().requires_grad_(True)
        (ref_upsample(yr.reshape(-1, 24), L_in).reshape(w.shape) * w).sum().backward()
        yg = y.clone().cuda().requires_grad_(True)
        (CU.upsample_pdf_right_inverse(yg, L_in) * w.cuda()).sum().backward()
        assert _close(yg.grad, yr.grad, 1e-4), (shape, L_in)
    # cdf_from_density (losses.py:11-12 differentiates it)
    p = torch.randn(6, 64, generator=g)
    w = torch.randn(6, 64, generator=g)
    pr = p.clone().requires_grad_(True)
    (ref_cdf(pr) * w).sum().backward()
    pg = p.clone().cuda().requires_grad_(True)
    (CU.cdf_from_density(pg) * w.cuda()).sum().backward()
    assert _close(pg.grad, pr.grad, 1e-4)
    # no_grad callers (trainer.py:284-296) are unaffected
    with torch.no_grad():
        assert not CU.upsample_pdf_right_inverse(torch.rand(2, 24, device="cuda"), 512).requires_grad


def _chain(ss, mix, up, z, gt, alpha, L):
    p = mix(ss(z, dim=1, eps=1e-6), alpha)
    x = up(p, L).clamp_min(0)
    x = x / x.sum(dim=1, keepdim=True).clamp_min(1e-6)
    return torch.nn.functional.l1_loss(x, gt), x


def test_restated_expressions_match_the_reference_gradients():
    """CPU: the expressions restated at the top of this file reproduce values AND gradients recorded from
    the unmodified reference (tests/golden/make_golden_autograd.py), so they are a valid yardstick."""
    g = np.load(GOLD)
    for k in range(int(g["n_chain"])):
        z = torch.from_numpy(g[f"chain{k}/z"]).requires_grad_(True)
        loss, x = _chain(ref_safe_softmax, ref_mix_with_uniform, ref_upsample, z, torch.from_numpy(g[f"chain{k}/gt"]),
                         float(g[f"chain{k}/alpha"]), int(g[f"chain{k}/L"]))
        loss.backward()
        assert _close(x, torch.from_numpy(g[f"chain{k}/x"]), 1e-5)
        assert _close(z.grad, torch.from_numpy(g[f"chain{k}/grad_z"]), 2e-4)
    p = torch.from_numpy(g["cdf/p"]).requires_grad_(True)
    (ref_cdf(p) * torch.from_numpy(g["cdf/w"])).sum().backward()
    assert _close(p.grad, torch.from_numpy(g["cdf/grad_p"]), 1e-5)


@pytest.mark.gpu
def test_library_gradients_match_the_reference_golden():
    """GPU: forward values and d loss / d logits of the mirrors against the gradients torch.autograd
    produced on the unmodified reference functions."""
    need_gpu()
    from attwarp_b200 import checkpoint_utils as CU, model as M
    g = np.load(GOLD)
    for k in range(int(g["n_chain"])):
        z = torch.from_numpy(g[f"chain{k}/z"]).cuda().requires_grad_(True)
        loss, x = _chain(M.safe_softmax, M.mix_with_uniform, CU.upsample_pdf_right_inverse, z,
                         torch.from_numpy(g[f"chain{k}/gt"]).cuda(), float(g[f"chain{k}/alpha"]), int(g[f"chain{k}/L"]))
        loss.backward()
        assert _close(x, torch.from_numpy(g[f"chain{k}/x"]), 1e-5)
        assert abs(float(loss.detach()) - float(g[f"chain{k}/loss"])) <= 1e-6
        assert _close(z.grad, torch.from_numpy(g[f"chain{k}/grad_z"]), 2e-4)
    p = torch.from_numpy(g["cdf/p"]).cuda().requires_grad_(True)
    Fp = CU.cdf_from_density(p)
    (Fp * torch.from_numpy(g["cdf/w"]).cuda()).sum().backward()
    assert _close(Fp, torch.from_numpy(g["cdf/F"]), 1e-5)
    assert _close(p.grad, torch.from_numpy(g["cdf/grad_p"]), 1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("alpha", [0.0, 0.1, 0.3])
def test_fused_softmax_mix_equals_the_two_calls(alpha):
    """model.safe_softmax_mix == mix_with_uniform(safe_softmax(.)) bit for bit (forward) and to float32 rounding
    (backward), incl. rows with NaN / +-inf logits."""
    need_gpu()
    from attwarp_b200 import model as M
    g = torch.Generator().manual_seed(17)
    z = torch.randn(128, 24, generator=g) * 3
    z[3, 5], z[4, 2], z[5, 1] = float("nan"), float("inf"), float("-inf")
    w = torch.randn(128, 24, generator=g).cuda()
    z1 = z.clone().cuda().requires_grad_(True)
    z2 = z.clone().cuda().requires_grad_(True)
    a = M.mix_with_uniform(M.safe_softmax(z1, dim=1, eps=1e-6), alpha)
    b = M.safe_softmax_mix(z2, alpha)
    assert torch.equal(a, b)
    (a * w).sum().backward()
    (b * w).sum().backward()
    assert _close(z2.grad, z1.grad, 1e-6)
