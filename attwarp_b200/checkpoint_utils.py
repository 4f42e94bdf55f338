"""Mirror of the warp helpers of ``model/marginalnet_full_dataset/checkpoint_utils.py``.

* ``cdf_from_density(p)``                               (checkpoint_utils.py:30-41)
* ``gt_marginals(A)``                                   (checkpoint_utils.py:43-51)
* ``upsample_pdf_right_inverse(y, target_len, eps)``    (checkpoint_utils.py:64-131)
* ``warp_from_cdf_torch(img, Fx_img, Fy_img, out_size)`` (checkpoint_utils.py:133-204)
* ``_make_strictly_increasing(Fcdf, eps)`` / ``resample_cdf(Fcdf, target_len)`` (checkpoint_utils.py:17-28, 53-62)
* ``adaptive_avg_pool2d_24`` -- the ``F.adaptive_avg_pool2d(A_full, (24, 24))`` prologue of
  ``trainer.py:197``

Same signatures and error behaviour.  The reference moves everything to the CPU and loops over
samples in Python around ``cv2.remap``; here tensors stay on (or are moved to) the GPU, the
batch is one launch per stage, and the result is returned on ``img.device`` with ``img.dtype``
exactly like checkpoint_utils.py:203.  No CPU fallback: a CUDA device is required.
The plotting helpers of the reference (checkpoint_utils.py:206-399) are out of scope.
"""

from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from . import ops
from ._lib import check, current_stream, load, ptr


def _cuda_device(t: torch.Tensor) -> torch.device:
    if t.is_cuda:
        return t.device
    if not torch.cuda.is_available():
        raise RuntimeError("attwarp_b200 needs a CUDA device (there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


class _CdfFromDensity(torch.autograd.Function):
    """Forward in the library; backward of clamp -> normalise -> cumsum -> last = 1 written out with torch
    tensor ops on the device (only the reference's unused l1_cdf_loss, losses.py:11-12, differentiates it)."""

    @staticmethod
    def forward(ctx, rows):
        lib = load()
        out = torch.empty_like(rows)
        with torch.cuda.device(rows.device):
            check(lib.attwarp_cdf_from_density(ptr(rows), rows.shape[0], rows.shape[1], ptr(out),
                                               current_stream(rows.device)))
        ctx.save_for_backward(rows)
        return out

    @staticmethod
    def backward(ctx, grad):
        (rows,) = ctx.saved_tensors
        live = torch.isfinite(rows) & (rows > 0)
        c = torch.where(live, rows, torch.zeros_like(rows))
        tot = c.sum(dim=1, keepdim=True)
        S = tot.clamp_min(1e-6)
        g = grad.clone().float()
        g[:, -1] = 0                                       # F[:, -1] = 1.0 overwrites the last element
        gq = g.flip(1).cumsum(1).flip(1)                   # d/dq of cumsum
        gc = gq / S - torch.where(tot > 1e-6, (gq * c).sum(dim=1, keepdim=True) / (S * S), torch.zeros_like(S))
        return torch.where(live, gc, torch.zeros_like(gc))


def cdf_from_density(p: torch.Tensor) -> torch.Tensor:
    """p: (B,N) -> (B,N) non-decreasing CDF in [0,1], ends at 1."""
    dev = _cuda_device(p)
    rows = p.to(dev).float().contiguous()
    assert rows.dim() == 2, "cdf_from_density expects (B, N)"
    return _CdfFromDensity.apply(rows).to(p.device)


def _make_strictly_increasing(Fcdf: torch.Tensor, eps: float = 1e-4) -> torch.Tensor:
    """(B,N) CDF rows -> strictly increasing rows in [0,1] ending at 1 (checkpoint_utils.py:17-28)."""
    lib = load()
    dev = _cuda_device(Fcdf)
    rows = Fcdf.detach().to(dev).float().contiguous()
    assert rows.dim() == 2, "_make_strictly_increasing expects (B, N)"
    out = torch.empty_like(rows)
    with torch.cuda.device(dev):
        check(lib.attwarp_make_strictly_increasing(ptr(rows), rows.shape[0], rows.shape[1], float(eps),
                                                   ptr(out), current_stream(dev)))
    return out.to(Fcdf.device)


def resample_cdf(Fcdf: torch.Tensor, target_len: int) -> torch.Tensor:
    """(B,N) CDF -> (B,target_len): strictly increasing -> linear interpolation (align_corners=True)
    -> strictly increasing (checkpoint_utils.py:53-62)."""
    lib = load()
    dev = _cuda_device(Fcdf)
    F1 = _make_strictly_increasing(Fcdf.detach().to(dev).float())
    B, N = F1.shape
    up = torch.empty(B, int(target_len), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.attwarp_interp_linear_rows(ptr(F1), B, N, int(target_len), ptr(up), current_stream(dev)))
    return _make_strictly_increasing(up).to(Fcdf.device)


def gt_marginals(A: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """A:(B,1,H,W) -> (px:(B,W), py:(B,H)) normalised."""
    lib = load()
    B, _, H, W = A.shape
    dev = _cuda_device(A)
    a = A.detach().to(dev).float()[:, 0].contiguous()
    px = torch.empty(B, W, dtype=torch.float32, device=dev)
    py = torch.empty(B, H, dtype=torch.float32, device=dev)
    wsb = lib.attwarp_maps_workspace_bytes(B, H, W)
    ws = ops._workspace(wsb, dev)
    with torch.cuda.device(dev):
        check(lib.attwarp_gt_marginals(ptr(a), B, H, W, ptr(ws), ws.numel(), ptr(px), ptr(py),
                                       current_stream(dev)))
    return px.to(A.device), py.to(A.device)


_M_cache = {}


def right_inverse_matrix(L_in: int, L_out: int, eps: float, device) -> torch.Tensor:
    """M = A^T (A A^T + eps I)^-1 as float32 [L_in, L_out] on ``device``; A is the
    AdaptiveAvgPool1d matrix with windows [floor(i*L_in/L_out), ceil((i+1)*L_in/L_out))
    (checkpoint_utils.py:104-121).  A constant of (L_in, L_out, eps): built once in float64 on
    the host and cached."""
    key = (int(L_in), int(L_out), float(eps), str(device))
    M = _M_cache.get(key)
    if M is None:
        i = np.arange(L_out, dtype=np.int64)
        starts = (i * L_in) // L_out
        ends = ((i + 1) * L_in + L_out - 1) // L_out
        A = np.zeros((L_out, L_in), dtype=np.float64)
        for k in range(L_out):
            A[k, starts[k]:ends[k]] = 1.0 / max(int(ends[k] - starts[k]), 1)
        G = A @ A.T
        if eps > 0:
            G = G + eps * np.eye(L_out)
        Mh = np.linalg.solve(G, A).T            # (L_in, L_out); G is symmetric
        M = torch.from_numpy(np.ascontiguousarray(Mh.astype(np.float32))).to(device)
        _M_cache[key] = M
    return M


class _UpsampleRightInverse(torch.autograd.Function):
    """rows [B, L_out] float32 (CUDA) -> rows M^T [B, L_in]; backward grad_x M (the op is linear)."""

    @staticmethod
    def forward(ctx, rows, M):
        lib = load()
        L_in, L_out = M.shape
        x = torch.empty(rows.shape[0], L_in, dtype=torch.float32, device=rows.device)
        with torch.cuda.device(rows.device):
            check(lib.attwarp_upsample_right_inverse(ptr(rows), ptr(M), rows.shape[0], L_out, L_in,
                                                     ptr(x), current_stream(rows.device)))
        ctx.save_for_backward(M)
        return x

    @staticmethod
    def backward(ctx, grad):
        (M,) = ctx.saved_tensors
        lib = load()
        L_in, L_out = M.shape
        g = grad.contiguous().float()
        gy = torch.empty(g.shape[0], L_out, dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            check(lib.attwarp_upsample_right_inverse_backward(ptr(g), ptr(M), g.shape[0], L_out, L_in, ptr(gy),
                                                              current_stream(g.device)))
        return gy, None


def upsample_pdf_right_inverse(y: torch.Tensor, target_len: int, eps: float = 1e-8) -> torch.Tensor:
    """Minimum-norm right inverse of AdaptiveAvgPool1d; y (L_out,), (B,L_out) or (B,C,L_out).
    Differentiable in y like the reference's (trainer.py:217-218 trains through it)."""
    if y.dim() not in (1, 2, 3):
        raise ValueError(f"upsample_pdf_right_inverse expects 1D/2D/3D y; got shape {tuple(y.shape)}")
    dev = _cuda_device(y)
    L_out, L_in = y.shape[-1], int(target_len)
    rows = y.to(dev).float().reshape(-1, L_out).contiguous()
    M = right_inverse_matrix(L_in, L_out, eps, dev)
    x = _UpsampleRightInverse.apply(rows, M)
    return x.reshape(tuple(y.shape[:-1]) + (L_in,)).to(device=y.device, dtype=y.dtype)


def adaptive_avg_pool2d_24(A: torch.Tensor, out_hw=(24, 24)) -> torch.Tensor:
    """F.adaptive_avg_pool2d(A, (24,24)) for A (B,1,H,W) float32 (trainer.py:197)."""
    lib = load()
    B, Cc, H, W = A.shape
    dev = _cuda_device(A)
    a = A.detach().to(dev).float().reshape(B * Cc, H, W).contiguous()
    gh, gw = out_hw
    out = torch.empty(B * Cc, gh, gw, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.attwarp_adaptive_avg_pool2d(ptr(a), B * Cc, H, W, gh, gw, ptr(out),
                                              current_stream(dev)))
    return out.reshape(B, Cc, gh, gw).to(A.device)


def pool_attention(A_full: torch.Tensor, transforms=None, out_hw=(24, 24)) -> torch.Tensor:
    """The trainer's attention prologue in one pass (trainer.py:172-197): ``A_full`` (B,1,H,W) is
    ``clamp_min(0)``-ed, samples whose entry of ``transforms`` is ``"sqrt"`` get a square root (the others --
    ``"iden"``, ``"none"`` -- stay as they are), and the result is pooled to ``out_hw``; the transformed
    full-resolution map is never written.  ``transforms=None`` is the plain ``adaptive_avg_pool2d`` the
    reference runs when the batch carries no dataset names."""
    if transforms is None:
        return adaptive_avg_pool2d_24(A_full, out_hw)
    lib = load()
    B, Cc, H, W = A_full.shape
    assert Cc == 1 and len(transforms) == B, "pool_attention expects (B,1,H,W) and one transform per sample"
    dev = _cuda_device(A_full)
    a = A_full.detach().to(dev).float().reshape(B, H, W).contiguous()
    mask = torch.tensor([1 if t == "sqrt" else 0 for t in transforms], dtype=torch.uint8, device=dev)
    gh, gw = out_hw
    out = torch.empty(B, gh, gw, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.attwarp_pool_attention(ptr(a), ptr(mask), B, H, W, gh, gw, ptr(out), current_stream(dev)))
    return out.reshape(B, 1, gh, gw).to(A_full.device)


def warp_from_cdf_torch(img: torch.Tensor, Fx_img: torch.Tensor, Fy_img: torch.Tensor,
                        out_size: tuple | None = None) -> torch.Tensor:
    """img (B,C,H,W) uint8 or float32; Fx_img (B,W), Fy_img (B,H) CDFs in [0,1];
    out_size (H_out, W_out) or None.  Returns (B,C,H_out,W_out) on img.device with img.dtype.

    Deliberate superset of the reference: C == 1 works (the reference crashes at
    checkpoint_utils.py:202 because cv2.remap drops the singleton channel)."""
    assert img.ndim == 4, f"img must be (B,C,H,W); got {img.shape}"
    B, Cc, H, W = img.shape
    H_out, W_out = (H, W) if out_size is None else out_size
    if Fx_img.shape[-1] != W:
        raise ValueError(f"Fx_img[0] length {Fx_img.shape[-1]} != image width W={W}")
    if Fy_img.shape[-1] != H:
        raise ValueError(f"Fy_img[0] length {Fy_img.shape[-1]} != image height H={H}")
    if img.dtype not in (torch.uint8, torch.float32):
        raise TypeError(f"warp_from_cdf_torch: img dtype {img.dtype} not supported (uint8/float32)")
    dev = _cuda_device(img)
    src = img.detach().to(dev).contiguous()
    Fx = Fx_img.detach().to(dev).float().reshape(B, W)
    Fy = Fy_img.detach().to(dev).float().reshape(B, H)
    map_x, map_y = ops.maps_from_cdf(Fx, Fy, (int(H_out), int(W_out)))
    out = ops.remap_bilinear(src, map_x, map_y, layout="chw")
    return out.to(device=img.device, dtype=img.dtype)
