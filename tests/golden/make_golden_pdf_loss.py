#!/usr/bin/env python
"""Golden loss values and GRADIENTS of the image-resolution PDF-L1 loss the reference trains with
(model/marginalnet_full_dataset/trainer.py:217-250), produced by torch.autograd on the UNMODIFIED reference
functions (build container only).

    python tests/golden/make_golden_pdf_loss.py        ->  tests/golden/pdf_loss.npz

Chain per axis (px_s = mix_with_uniform(safe_softmax(z)), px_gt = gt_marginals(A) with A the 24x24 pooled attention):
    px_img    = upsample_pdf_right_inverse(px_s, W).clamp_min(0)
    px_gt_img = upsample_pdf_right_inverse(px_gt, W).clamp_min(0)
    px_img, px_gt_img /= their row sums .clamp_min(1e-6)
    L_pdf = F.l1_loss(px_img, px_gt_img) + F.l1_loss(py_img, py_gt_img)
Stored per case: px_s, py_s, px_gt, py_gt, W, H, L_pdf and d L_pdf / d px_s, d py_s.  One case overrides some GT rows
with the uniform PDF (trainer.py:236-239: samples whose transform is 'none') and one feeds PDFs with negative
up-sampled bins (non-divisible lengths), which exercises the clamp.
"""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

CASES = [dict(B=16, W=512, H=512, alpha=0.0, scale=2.0, none_rows=()),
         dict(B=16, W=336, H=500, alpha=0.1, scale=3.0, none_rows=(1, 5)),
         dict(B=7, W=100, H=64, alpha=0.05, scale=6.0, none_rows=()),
         dict(B=4, W=1344, H=224, alpha=0.0, scale=1.0, none_rows=(0,))]


def main():
    import torch
    import torch.nn.functional as F

    from oracle import ref_loader as R
    assert R.available(), "reference tree not found"
    cu, mm = R.checkpoint_utils(), R.marginalnet_model()
    out = {}
    for k, c in enumerate(CASES):
        g = torch.Generator().manual_seed(5100 + k)
        B, W, H = c["B"], c["W"], c["H"]
        zx, zy = torch.randn(B, 24, generator=g) * c["scale"], torch.randn(B, 24, generator=g) * c["scale"]
        px_s = mm.mix_with_uniform(mm.safe_softmax(zx), c["alpha"]).detach().requires_grad_(True)
        py_s = mm.mix_with_uniform(mm.safe_softmax(zy), c["alpha"]).detach().requires_grad_(True)
        A = torch.rand(B, 1, 24, 24, generator=g) ** 3
        px_gt, py_gt = cu.gt_marginals(A)
        for r in c["none_rows"]:
            px_gt[r] = 1.0 / 24
            py_gt[r] = 1.0 / 24
        px_img = cu.upsample_pdf_right_inverse(px_s, W).clamp_min(0)
        py_img = cu.upsample_pdf_right_inverse(py_s, H).clamp_min(0)
        px_gt_img = cu.upsample_pdf_right_inverse(px_gt, W).clamp_min(0)
        py_gt_img = cu.upsample_pdf_right_inverse(py_gt, H).clamp_min(0)
        px_img = px_img / px_img.sum(dim=1, keepdim=True).clamp_min(1e-6)
        py_img = py_img / py_img.sum(dim=1, keepdim=True).clamp_min(1e-6)
        px_gt_img = px_gt_img / px_gt_img.sum(dim=1, keepdim=True).clamp_min(1e-6)
        py_gt_img = py_gt_img / py_gt_img.sum(dim=1, keepdim=True).clamp_min(1e-6)
        L_pdf = F.l1_loss(px_img, px_gt_img) + F.l1_loss(py_img, py_gt_img)
        (3.0 * L_pdf).backward()                      # an upstream gradient other than 1 (cfg.w_cdf, AMP scaling)
        out[f"case{k}/px_s"], out[f"case{k}/py_s"] = px_s.detach().numpy(), py_s.detach().numpy()
        out[f"case{k}/px_gt"], out[f"case{k}/py_gt"] = px_gt.numpy(), py_gt.numpy()
        out[f"case{k}/W"], out[f"case{k}/H"] = np.int64(W), np.int64(H)
        out[f"case{k}/loss"] = np.float64(L_pdf.item())
        out[f"case{k}/upstream"] = np.float64(3.0)
        out[f"case{k}/grad_px_s"], out[f"case{k}/grad_py_s"] = px_s.grad.numpy(), py_s.grad.numpy()
        out[f"case{k}/neg_bins"] = np.int64(int((cu.upsample_pdf_right_inverse(px_s.detach(), W) < 0).sum()))
    out["n_cases"] = np.int64(len(CASES))
    path = os.path.join(HERE, "pdf_loss.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", [int(out[f"case{k}/neg_bins"]) for k in range(len(CASES))],
          [float(out[f"case{k}/loss"]) for k in range(len(CASES))])


if __name__ == "__main__":
    main()
