"""Mirror of the PDF-normalisation helpers of ``model/marginalnet_full_dataset/model.py``.

* ``safe_softmax(logits, dim=1, eps=1e-6)``   (model.py:8-14)
* ``mix_with_uniform(p, alpha)``              (model.py:98-101)

CUDA tensors only; computed by libattwarp_sm100.so (no CPU fallback).  The MarginalNet network
itself (dense convolutions, model.py:17-95) is out of scope.
"""

from __future__ import annotations

import torch

from ._lib import check, current_stream, load, ptr, require_cuda


def _rows(t: torch.Tensor, dim: int):
    """View ``t`` as [B, N] float32 rows along ``dim``; returns (rows, restore_fn)."""
    dim = dim % t.dim()
    moved = t.movedim(dim, -1)
    shape = moved.shape
    rows = moved.reshape(-1, shape[-1]).contiguous().float()

    def restore(r):
        return r.reshape(shape).movedim(-1, dim).to(t.dtype)

    return rows, restore


def safe_softmax(logits: torch.Tensor, dim: int = 1, eps: float = 1e-6) -> torch.Tensor:
    lib = load()
    require_cuda(logits)
    rows, restore = _rows(logits, dim)
    out = torch.empty_like(rows)
    with torch.cuda.device(rows.device):
        check(lib.attwarp_safe_softmax(ptr(rows), rows.shape[0], rows.shape[1], float(eps),
                                       ptr(out), current_stream(rows.device)))
    return restore(out)


def mix_with_uniform(p: torch.Tensor, alpha: float) -> torch.Tensor:
    if alpha <= 0:                       # model.py:99-100 returns the input itself
        return p
    lib = load()
    require_cuda(p)
    assert p.dim() == 2, "mix_with_uniform expects (B, N)"
    rows = p.contiguous().float()
    out = torch.empty_like(rows)
    with torch.cuda.device(rows.device):
        check(lib.attwarp_mix_with_uniform(ptr(rows), rows.shape[0], rows.shape[1], float(alpha),
                                           ptr(out), current_stream(rows.device)))
    return out.to(p.dtype)
