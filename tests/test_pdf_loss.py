"""The fused image-resolution PDF-L1 loss (mnfd/trainer.py:217-250) against loss values
and gradients torch.autograd produced on the UNMODIFIED reference chain (tests/golden/pdf_loss.npz, written by
tests/golden/make_golden_pdf_loss.py).  CPU: the oracle's forward restatement.  GPU: forward <= 1e-5 relative,
gradients <= 1e-5 of the largest gradient entry (the gradient of an L1 loss is a sum of +-1/(B L) terms: entries
cancel to ~0, so the bar is relative to the scale), one launch each way."""

import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import torch_path as OT


@pytest.fixture(scope="module")
def gl():
    return np.load(os.path.join(GOLDEN_DIR, "pdf_loss.npz"))


def _cases(gl):
    return range(int(gl["n_cases"]))


def test_oracle_pdf_l1_loss_forward(gl):
    for k in _cases(gl):
        got = OT.pdf_l1_loss(gl[f"case{k}/px_s"], gl[f"case{k}/py_s"], gl[f"case{k}/px_gt"], gl[f"case{k}/py_gt"],
                             (int(gl[f"case{k}/H"]), int(gl[f"case{k}/W"])))
        assert abs(got - float(gl[f"case{k}/loss"])) <= 1e-5 * float(gl[f"case{k}/loss"])


@pytest.mark.gpu
def test_gpu_pdf_l1_loss_forward_backward(gl):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from attwarp_b200 import trainer as T
    for k in _cases(gl):
        px = torch.from_numpy(gl[f"case{k}/px_s"]).cuda().requires_grad_(True)
        py = torch.from_numpy(gl[f"case{k}/py_s"]).cuda().requires_grad_(True)
        gx, gy = torch.from_numpy(gl[f"case{k}/px_gt"]).cuda(), torch.from_numpy(gl[f"case{k}/py_gt"]).cuda()
        hw = (int(gl[f"case{k}/H"]), int(gl[f"case{k}/W"]))
        loss = T.pdf_l1_loss(px, py, gx, gy, hw)
        ref = float(gl[f"case{k}/loss"])
        assert abs(loss.item() - ref) <= 1e-5 * ref, (k, loss.item(), ref)
        (float(gl[f"case{k}/upstream"]) * loss).backward()
        for got, want in ((px.grad, gl[f"case{k}/grad_px_s"]), (py.grad, gl[f"case{k}/grad_py_s"])):
            scale = np.abs(want).max()
            assert np.abs(got.cpu().numpy() - want).max() <= 1e-5 * scale + 1e-9, (k, np.abs(got.cpu().numpy() - want).max(), scale)
        # a second evaluation re-uses nothing stale (the kernel leaves its workspace ready) and is deterministic
        assert T.pdf_l1_loss(px.detach(), py.detach(), gx, gy, hw).item() == loss.item()


@pytest.mark.gpu
def test_gpu_pdf_l1_loss_matches_unfused_chain():
    """Against the package's own unfused mirrors (upsample_pdf_right_inverse autograd Function + torch ops), at the
    trainer's shapes (B=128, 512^2)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import torch.nn.functional as F
    from attwarp_b200 import checkpoint_utils as CU, trainer as T
    g = torch.Generator().manual_seed(3)
    B, W, H = 128, 512, 512
    px = torch.softmax(torch.randn(B, 24, generator=g) * 2, -1).cuda().requires_grad_(True)
    py = torch.softmax(torch.randn(B, 24, generator=g) * 2, -1).cuda().requires_grad_(True)
    gx = torch.softmax(torch.randn(B, 24, generator=g), -1).cuda()
    gy = torch.softmax(torch.randn(B, 24, generator=g), -1).cuda()

    def norm(t):
        return t / t.sum(dim=1, keepdim=True).clamp_min(1e-6)

    ref = F.l1_loss(norm(CU.upsample_pdf_right_inverse(px, W).clamp_min(0)), norm(CU.upsample_pdf_right_inverse(gx, W).clamp_min(0))) + \
        F.l1_loss(norm(CU.upsample_pdf_right_inverse(py, H).clamp_min(0)), norm(CU.upsample_pdf_right_inverse(gy, H).clamp_min(0)))
    ref.backward()
    rgx, rgy = px.grad.clone(), py.grad.clone()
    px.grad = py.grad = None
    loss = T.pdf_l1_loss(px, py, gx, gy, (H, W))
    loss.backward()
    assert abs(loss.item() - ref.item()) <= 1e-5 * ref.item()
    for got, want in ((px.grad, rgx), (py.grad, rgy)):
        assert (got - want).abs().max().item() <= 1e-5 * want.abs().max().item() + 1e-9
