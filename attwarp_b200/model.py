"""Mirror of the PDF-normalisation helpers of ``model/marginalnet_full_dataset/model.py``.

* ``safe_softmax(logits, dim=1, eps=1e-6)``   (model.py:8-14)
* ``mix_with_uniform(p, alpha)``              (model.py:98-101)
* ``safe_softmax_mix(logits, alpha)``         the two back to back in one launch (MarginalNet.forward's last op,
  model.py:93-94, followed by trainer.py:212-214); bit-identical to calling them in turn

CUDA tensors only; computed by libattwarp_sm100.so (no CPU fallback).  The MarginalNet network
itself (dense convolutions, model.py:17-95) is out of scope.
"""

from __future__ import annotations

import torch

from ._lib import check, current_stream, load, ptr, require_cuda


def _rows(t: torch.Tensor, dim: int):
    """View ``t`` as [B, N] float32 rows along ``dim``; returns (rows, restore_fn)."""
    dim = dim % t.dim()
    if t.dim() == 2 and dim == 1 and t.dtype == torch.float32 and t.is_contiguous():
        return t, lambda r: r                       # the trainer's case: no views, no copies
    moved = t.movedim(dim, -1)
    shape = moved.shape
    rows = moved.reshape(-1, shape[-1]).contiguous().float()

    def restore(r):
        return r.reshape(shape).movedim(-1, dim).to(t.dtype)

    return rows, restore


class _SafeSoftmaxRows(torch.autograd.Function):
    """rows [B, N] float32 -> safe_softmax rows; backward = what autograd gives model.py:8-14."""

    @staticmethod
    def forward(ctx, rows, eps):
        lib = load()
        out = torch.empty_like(rows)
        with torch.cuda.device(rows.device):
            check(lib.attwarp_safe_softmax(ptr(rows), rows.shape[0], rows.shape[1], float(eps),
                                           ptr(out), current_stream(rows.device)))
        ctx.save_for_backward(rows)
        ctx.eps = float(eps)
        return out

    @staticmethod
    def backward(ctx, grad):
        (rows,) = ctx.saved_tensors
        lib = load()
        g = grad.contiguous().float()
        gz = torch.empty_like(rows)
        with torch.cuda.device(rows.device):
            check(lib.attwarp_safe_softmax_backward(ptr(rows), ptr(g), rows.shape[0], rows.shape[1], ctx.eps,
                                                    ptr(gz), current_stream(rows.device)))
        return gz, None


class _MixWithUniform(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rows, alpha):
        lib = load()
        out = torch.empty_like(rows)
        with torch.cuda.device(rows.device):
            check(lib.attwarp_mix_with_uniform(ptr(rows), rows.shape[0], rows.shape[1], float(alpha),
                                               ptr(out), current_stream(rows.device)))
        ctx.alpha = float(alpha)
        return out

    @staticmethod
    def backward(ctx, grad):
        lib = load()
        g = grad.contiguous().float()
        gp = torch.empty_like(g)
        with torch.cuda.device(g.device):
            check(lib.attwarp_mix_with_uniform_backward(ptr(g), g.shape[0], g.shape[1], ctx.alpha, ptr(gp),
                                                        current_stream(g.device)))
        return gp, None


def safe_softmax(logits: torch.Tensor, dim: int = 1, eps: float = 1e-6) -> torch.Tensor:
    """Differentiable like the reference's (MarginalNet.forward ends in it, model.py:93-94)."""
    require_cuda(logits)
    rows, restore = _rows(logits, dim)
    return restore(_SafeSoftmaxRows.apply(rows, eps))


def mix_with_uniform(p: torch.Tensor, alpha: float) -> torch.Tensor:
    """Differentiable like the reference's (trainer.py:213-214 trains through it)."""
    if alpha <= 0:                       # model.py:99-100 returns the input itself
        return p
    require_cuda(p)
    assert p.dim() == 2, "mix_with_uniform expects (B, N)"
    return _MixWithUniform.apply(p.contiguous().float(), alpha).to(p.dtype)


class _SafeSoftmaxMixRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rows, eps, alpha):
        lib = load()
        out = torch.empty_like(rows)
        with torch.cuda.device(rows.device):
            check(lib.attwarp_safe_softmax_mix(ptr(rows), rows.shape[0], rows.shape[1], float(eps), float(alpha),
                                               ptr(out), current_stream(rows.device)))
        ctx.save_for_backward(rows)
        ctx.eps, ctx.alpha = float(eps), float(alpha)
        return out

    @staticmethod
    def backward(ctx, grad):
        (rows,) = ctx.saved_tensors
        lib = load()
        g = grad.contiguous().float()
        gz = torch.empty_like(rows)
        with torch.cuda.device(rows.device):
            check(lib.attwarp_safe_softmax_mix_backward(ptr(rows), ptr(g), rows.shape[0], rows.shape[1], ctx.eps,
                                                        ctx.alpha, ptr(gz), current_stream(rows.device)))
        return gz, None, None


def safe_softmax_mix(logits: torch.Tensor, alpha: float, dim: int = 1, eps: float = 1e-6) -> torch.Tensor:
    """``mix_with_uniform(safe_softmax(logits, dim, eps), alpha)`` in one launch (and one backward launch)."""
    require_cuda(logits)
    rows, restore = _rows(logits, dim)
    return restore(_SafeSoftmaxMixRows.apply(rows, eps, alpha))
