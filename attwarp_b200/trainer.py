"""Fused pieces of the MarginalNet training step (``model/marginalnet_full_dataset/trainer.py``) -- SURVEY.md
section 8(f) N4.  The trainer itself (optimizer, AMP, data, logging) is out of scope; these are the parts of its
step that sit on the warp path and that the reference runs as long chains of small torch kernels:

* ``pdf_l1_loss(px_s, py_s, px_gt, py_gt, image_hw)``   trainer.py:217-250 -- up-sample the predicted and the
  ground-truth PDFs to image resolution, clamp, renormalise, L1 -- ONE launch forward, ONE backward
  (``torch.autograd.Function``; the ground truth carries no gradient, like ``gt_marginals`` of the data);
* ``pool_attention``                                    trainer.py:186-197 (re-exported from checkpoint_utils).

Swapping these in keeps training numerically equivalent: tests/test_gpu_autograd.py and tests/test_pdf_loss.py hold
them to gradients recorded from the unmodified reference.
"""

from __future__ import annotations

import torch

from ._lib import check, current_stream, load, ptr
from .checkpoint_utils import _cuda_device, pool_attention, right_inverse_matrix  # noqa: F401


class _PdfL1Loss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, px, py, gx, gy, W, H, eps):
        lib = load()
        dev = px.device
        B = px.shape[0]
        mats = (right_inverse_matrix(W, px.shape[1], eps, dev), right_inverse_matrix(H, py.shape[1], eps, dev),
                right_inverse_matrix(W, gx.shape[1], eps, dev), right_inverse_matrix(H, gy.shape[1], eps, dev))
        ws = torch.zeros(int(lib.attwarp_pdf_l1_loss_workspace_bytes(B)), dtype=torch.uint8, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.attwarp_pdf_l1_loss(ptr(px), ptr(py), ptr(gx), ptr(gy), B, px.shape[1], py.shape[1], gx.shape[1],
                                          gy.shape[1], ptr(mats[0]), ptr(mats[1]), ptr(mats[2]), ptr(mats[3]), W, H,
                                          ptr(ws), ws.numel(), ptr(loss), current_stream(dev)))
        ctx.save_for_backward(px, py, gx, gy, *mats)
        ctx.dims = (W, H)
        return loss

    @staticmethod
    def backward(ctx, grad):
        px, py, gx, gy, Mx, My, Mgx, Mgy = ctx.saved_tensors
        lib = load()
        W, H = ctx.dims
        dev = px.device
        up = grad.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        dpx, dpy = torch.empty_like(px), torch.empty_like(py)
        with torch.cuda.device(dev):
            check(lib.attwarp_pdf_l1_loss_backward(ptr(px), ptr(py), ptr(gx), ptr(gy), px.shape[0], px.shape[1],
                                                   py.shape[1], gx.shape[1], gy.shape[1], ptr(Mx), ptr(My), ptr(Mgx),
                                                   ptr(Mgy), W, H, ptr(up), ptr(dpx), ptr(dpy), current_stream(dev)))
        return dpx, dpy, None, None, None, None, None


def pdf_l1_loss(px_s: torch.Tensor, py_s: torch.Tensor, px_gt: torch.Tensor, py_gt: torch.Tensor, image_hw,
                eps: float = 1e-8) -> torch.Tensor:
    """``L_pdf`` of trainer.py:217-250: px_s (B,Nx), py_s (B,Ny) predicted PDFs (after ``mix_with_uniform``), px_gt
    (B,Ngx), py_gt (B,Ngy) ``gt_marginals`` of the pooled attention, ``image_hw = (img.size(-2), img.size(-1))``.
    Returns the scalar ``F.l1_loss(px_img, px_gt_img) + F.l1_loss(py_img, py_gt_img)``, differentiable in px_s and
    py_s."""
    dev = _cuda_device(px_s)
    H, W = int(image_hw[0]), int(image_hw[1])
    if px_s.dim() != 2 or py_s.dim() != 2 or px_gt.dim() != 2 or py_gt.dim() != 2:
        raise ValueError("pdf_l1_loss expects (B, N) PDFs")
    B = px_s.shape[0]
    if py_s.shape[0] != B or px_gt.shape[0] != B or py_gt.shape[0] != B:
        raise ValueError("pdf_l1_loss: batch sizes differ")
    px = px_s.to(dev).float().contiguous()
    py = py_s.to(dev).float().contiguous()
    gx = px_gt.detach().to(dev).float().contiguous()
    gy = py_gt.detach().to(dev).float().contiguous()
    return _PdfL1Loss.apply(px, py, gx, gy, W, H, float(eps)).to(px_s.device)
