#!/usr/bin/env python
"""Golden GRADIENTS for the helpers the reference trains through, produced by torch.autograd on the
UNMODIFIED REFERENCE functions (build container only).

    python tests/golden/make_golden_autograd.py        ->  tests/golden/autograd.npz

Chain (model/marginalnet_full_dataset/trainer.py:209-250, model.py:93-94):
    p     = safe_softmax(z, dim=1, eps=1e-6)
    p_s   = mix_with_uniform(p, alpha)
    x     = upsample_pdf_right_inverse(p_s, L).clamp_min(0);  x = x / x.sum(1, keepdim).clamp_min(1e-6)
    loss  = F.l1_loss(x, gt)
Stored per case: z, gt, alpha, L, the forward x, the loss and d loss / d z; plus one case of
cdf_from_density (losses.py:11-12): p, weights w, F and d (F * w).sum() / d p.
"""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

CASES = [(0.0, 512), (0.1, 512), (0.3, 336), (0.05, 100)]


def main():
    import torch

    from oracle import ref_loader as R
    assert R.available(), "reference tree not found"
    cu, mm = R.checkpoint_utils(), R.marginalnet_model()
    out = {}
    for k, (alpha, L) in enumerate(CASES):
        g = torch.Generator().manual_seed(4100 + k)
        z = (torch.randn(16, 24, generator=g) * 2).requires_grad_(True)
        gt = torch.softmax(torch.randn(16, L, generator=g), -1)
        p = mm.safe_softmax(z, dim=1, eps=1e-6)
        ps = mm.mix_with_uniform(p, alpha)
        x = cu.upsample_pdf_right_inverse(ps, L).clamp_min(0)
        x = x / x.sum(dim=1, keepdim=True).clamp_min(1e-6)
        loss = torch.nn.functional.l1_loss(x, gt)
        loss.backward()
        out[f"chain{k}/z"] = z.detach().numpy()
        out[f"chain{k}/gt"] = gt.numpy()
        out[f"chain{k}/alpha"] = np.float64(alpha)
        out[f"chain{k}/L"] = np.int64(L)
        out[f"chain{k}/x"] = x.detach().numpy()
        out[f"chain{k}/loss"] = np.float64(loss.item())
        out[f"chain{k}/grad_z"] = z.grad.numpy()
    g = torch.Generator().manual_seed(4200)
    p = torch.randn(6, 64, generator=g).requires_grad_(True)
    w = torch.randn(6, 64, generator=g)
    Fp = cu.cdf_from_density(p)
    (Fp * w).sum().backward()
    out["cdf/p"], out["cdf/w"] = p.detach().numpy(), w.numpy()
    out["cdf/F"], out["cdf/grad_p"] = Fp.detach().numpy(), p.grad.numpy()
    out["n_chain"] = np.int64(len(CASES))
    path = os.path.join(HERE, "autograd.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
