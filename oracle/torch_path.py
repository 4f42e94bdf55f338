"""Oracle: NumPy restatement of the reference's torch-side warp helpers.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Follows ``/root/reference/model/marginalnet_full_dataset/``:

* ``_make_strictly_increasing`` ........ ``checkpoint_utils.py:17-28``
* ``cdf_from_density`` ................. ``checkpoint_utils.py:30-41``
* ``gt_marginals`` ..................... ``checkpoint_utils.py:43-51``
* ``resample_cdf`` ..................... ``checkpoint_utils.py:53-62``
* ``upsample_pdf_right_inverse`` ....... ``checkpoint_utils.py:64-131``
* ``warp_from_cdf_torch`` .............. ``checkpoint_utils.py:133-204``
* ``safe_softmax`` / ``mix_with_uniform`` ``model.py:8-14`` / ``model.py:98-101``
* ``adaptive_avg_pool2d -> (24,24)`` .... ``trainer.py:197`` (PyTorch window rule
  ``[floor(i*N/g), ceil((i+1)*N/g))``)

All arrays are float32 where the reference computes in float32.  ``torch.cumsum`` on CPU
accumulates float32 rows in float64 and rounds every output element to float32 (verified in
this container against torch 2.11), which is what ``cumsum_f32`` restates.
"""

from __future__ import annotations

import numpy as np

from . import numpy_path as _np_path

F32 = np.float32


def _nan_to_num(x, nan=0.0, posinf=None, neginf=None):
    x = np.array(x, dtype=F32, copy=True)
    fin = np.finfo(F32)
    x[np.isnan(x)] = nan
    x[np.isposinf(x)] = fin.max if posinf is None else posinf
    x[np.isneginf(x)] = fin.min if neginf is None else neginf
    return x


def cumsum_f32(p):
    return np.cumsum(np.asarray(p, dtype=F32).astype(np.float64), axis=-1).astype(F32)


def safe_softmax(logits, eps=1e-6):
    """(B,N) logits -> (B,N) probabilities; model.py:8-14 with dim=1."""
    z = _nan_to_num(logits, 0.0, 0.0, 0.0)
    z = z - z.max(axis=1, keepdims=True)
    e = np.exp(z.astype(F32)).astype(F32)
    p = (e / e.sum(axis=1, keepdims=True, dtype=F32)).astype(F32)
    p = _nan_to_num(p, 0.0, 0.0, 0.0)
    return (p / np.maximum(p.sum(axis=1, keepdims=True, dtype=F32), F32(eps))).astype(F32)


def mix_with_uniform(p, alpha):
    """model.py:98-101."""
    p = np.asarray(p, dtype=F32)
    if alpha <= 0:
        return p
    return (F32(1 - alpha) * p + F32(alpha / p.shape[1])).astype(F32)


def cdf_from_density(p):
    """(B,N) -> (B,N); checkpoint_utils.py:30-41."""
    p = np.asarray(p, dtype=F32)
    q = np.where(np.isnan(p), p, np.maximum(p, F32(0)))     # clamp_min propagates NaN
    q = _nan_to_num(q, 0.0, 0.0, 0.0)
    denom = np.maximum(q.sum(axis=1, keepdims=True, dtype=F32), F32(1e-6))
    q = (q / denom).astype(F32)
    Fp = cumsum_f32(q)
    Fp[:, -1] = 1.0
    return Fp


def gt_marginals(A):
    """(B,1,H,W) -> ((B,W), (B,H)); checkpoint_utils.py:43-51."""
    A = np.asarray(A, dtype=F32)
    Apos = np.maximum(A, F32(0))[:, 0]
    mx = Apos.sum(axis=1, dtype=F32)          # sum over H -> (B,W)
    my = Apos.sum(axis=2, dtype=F32)          # sum over W -> (B,H)
    mx = mx / np.maximum(mx.sum(axis=1, keepdims=True, dtype=F32), F32(1e-6))
    my = my / np.maximum(my.sum(axis=1, keepdims=True, dtype=F32), F32(1e-6))
    return mx.astype(F32), my.astype(F32)


def pool_windows(L_in, L_out):
    """AdaptiveAvgPool1d windows: start=floor(i*L_in/L_out), end=ceil((i+1)*L_in/L_out)."""
    i = np.arange(L_out, dtype=np.int64)
    starts = (i * L_in) // L_out
    ends = ((i + 1) * L_in + L_out - 1) // L_out
    return starts, ends


def pooling_matrix(L_in, L_out, dtype=F32):
    starts, ends = pool_windows(L_in, L_out)
    A = np.zeros((L_out, L_in), dtype=dtype)
    for k in range(L_out):
        A[k, starts[k]:ends[k]] = 1.0 / max(int(ends[k] - starts[k]), 1)
    return A


def right_inverse_matrix(L_in, L_out, eps=1e-8, dtype=np.float64):
    """M = A^T (A A^T + eps I)^-1, shape (L_in, L_out); checkpoint_utils.py:104-121."""
    A = pooling_matrix(L_in, L_out, dtype)
    G = A @ A.T
    if eps > 0:
        G = G + dtype(eps) * np.eye(L_out, dtype=dtype)
    return (A.T @ np.linalg.inv(G)).astype(dtype)


def upsample_pdf_right_inverse(y, target_len, eps=1e-8):
    """y (..., L_out) float32 -> (..., target_len); checkpoint_utils.py:64-131 (solve in fp32)."""
    y = np.asarray(y, dtype=F32)
    if y.ndim > 3:
        raise ValueError(f"upsample_pdf_right_inverse expects 1D/2D/3D y; got shape {y.shape}")
    yN = y.reshape(-1, y.shape[-1])
    L_out, L_in = yN.shape[1], int(target_len)
    A = pooling_matrix(L_in, L_out, F32)
    G = (A @ A.T).astype(F32)
    if eps > 0:
        G = (G + F32(eps) * np.eye(L_out, dtype=F32)).astype(F32)
    tmp = np.linalg.solve(G, yN.T).astype(F32)
    x = (A.T @ tmp).T.astype(F32)
    return x.reshape(y.shape[:-1] + (L_in,))


def adaptive_avg_pool2d(A, out_hw=(24, 24)):
    """(B,1,H,W) float32 -> (B,1,gh,gw) window means (PyTorch adaptive pooling windows)."""
    A = np.asarray(A, dtype=F32)
    B, C, H, W = A.shape
    gh, gw = out_hw
    ys, ye = pool_windows(H, gh)
    xs, xe = pool_windows(W, gw)
    out = np.empty((B, C, gh, gw), dtype=F32)
    for i in range(gh):
        for j in range(gw):
            win = A[:, :, ys[i]:ye[i], xs[j]:xe[j]].astype(np.float64)
            out[:, :, i, j] = win.mean(axis=(2, 3))
    return out


def make_strictly_increasing(Fcdf, eps=1e-4):
    """checkpoint_utils.py:17-28."""
    Fc = _nan_to_num(Fcdf, 0.0, 1.0, 0.0)
    Fnd = np.maximum.accumulate(Fc, axis=1)
    B, N = Fnd.shape
    min_step = F32(eps / max(N, 1))
    d = np.maximum(Fnd[:, 1:] - Fnd[:, :-1], min_step).astype(F32)
    Ffix = np.concatenate([Fnd[:, :1], Fnd[:, :1] + cumsum_f32(d)], axis=1).astype(F32)
    last = np.maximum(Ffix[:, -1:], F32(1e-6))
    Ffix = np.clip((Ffix / last).astype(F32), F32(0), F32(1))
    Ffix[:, -1] = 1.0
    return Ffix


def interpolate_linear_align_corners(F, target_len):
    """F.interpolate(mode='linear', align_corners=True) on (B,N) float32 rows."""
    F = np.asarray(F, dtype=F32)
    B, N = F.shape
    L = int(target_len)
    scale = F32((N - 1) / (L - 1)) if L > 1 else F32(0)
    pos = (scale * np.arange(L, dtype=F32)).astype(F32)
    i0 = np.minimum(pos.astype(np.int64), N - 1)
    i1 = np.minimum(i0 + 1, N - 1)
    lam1 = (pos - i0.astype(F32)).astype(F32)
    lam0 = (F32(1) - lam1).astype(F32)
    return (lam0 * F[:, i0] + lam1 * F[:, i1]).astype(F32)


def resample_cdf(Fcdf, target_len):
    """checkpoint_utils.py:53-62."""
    F1 = make_strictly_increasing(np.asarray(Fcdf, dtype=F32))
    return make_strictly_increasing(interpolate_linear_align_corners(F1, target_len))


def knots_from_cdf(F_row, n_out):
    """checkpoint_utils.py:171-184: float64 knots incl. the tie-break branch.
    Returns (xp, tie_break_fired)."""
    F_row = np.asarray(F_row, dtype=F32).reshape(-1)
    xp = np.concatenate(([0.0], F_row.astype(np.float64))) * float(n_out)
    xp[-1] = n_out
    fired = bool(np.any(np.diff(xp) <= 0))
    if fired:
        inc = F32(1e-4 / max(n_out, 1)) * np.arange(xp.size, dtype=F32)   # float32 product
        xp = xp + inc.astype(np.float64)
    return xp, fired


def maps_from_cdf(Fx, Fy, out_size):
    """Stage 4 of the torch path. Fx (B,W), Fy (B,H) -> float32 (B,W_out), (B,H_out)."""
    H_out, W_out = out_size
    B = Fx.shape[0]
    mx = np.empty((B, W_out), dtype=F32)
    my = np.empty((B, H_out), dtype=F32)
    for b in range(B):
        xp, _ = knots_from_cdf(Fx[b], W_out)
        yp, _ = knots_from_cdf(Fy[b], H_out)
        mx[b] = _np_path.interp_restated(np.arange(W_out, dtype=F32), xp).astype(F32)
        my[b] = _np_path.interp_restated(np.arange(H_out, dtype=F32), yp).astype(F32)
    return mx, my


def warp_from_cdf(img, Fx, Fy, out_size=None, remap_backend="restated"):
    """img (B,C,H,W) uint8|float32, Fx (B,W), Fy (B,H) -> (B,C,H_out,W_out);
    checkpoint_utils.py:133-204."""
    img = np.asarray(img)
    assert img.ndim == 4, f"img must be (B,C,H,W); got {img.shape}"
    B, C, H, W = img.shape
    H_out, W_out = (H, W) if out_size is None else out_size
    if Fx.shape[1] != W:
        raise ValueError(f"Fx_img length {Fx.shape[1]} != image width W={W}")
    if Fy.shape[1] != H:
        raise ValueError(f"Fy_img length {Fy.shape[1]} != image height H={H}")
    mx, my = maps_from_cdf(np.asarray(Fx), np.asarray(Fy), (H_out, W_out))
    fn = _np_path.remap if remap_backend == "restated" else _np_path.remap_cv2
    out = np.empty((B, C, H_out, W_out), dtype=img.dtype)
    for b in range(B):
        hwc = np.ascontiguousarray(np.transpose(img[b], (1, 2, 0)))
        w = fn(hwc, mx[b], my[b])
        if w.ndim == 2:
            w = w[..., None]
        out[b] = np.transpose(w, (2, 0, 1))
    return out


def pdf_l1_loss(px_s, py_s, px_gt, py_gt, image_hw, eps=1e-8):
    """``L_pdf`` of trainer.py:217-250 (forward value): up-sample predicted and ground-truth PDFs to image
    resolution, ``clamp_min(0)``, divide by the row sum ``clamp_min(1e-6)``, mean absolute difference per axis."""
    H, W = image_hw
    tot = 0.0
    for p, g, L in ((px_s, px_gt, W), (py_s, py_gt, H)):
        a = np.maximum(upsample_pdf_right_inverse(np.asarray(p, F32), L, eps), 0).astype(F32)
        b = np.maximum(upsample_pdf_right_inverse(np.asarray(g, F32), L, eps), 0).astype(F32)
        a = a / np.maximum(a.sum(axis=1, keepdims=True, dtype=F32), F32(1e-6))
        b = b / np.maximum(b.sum(axis=1, keepdims=True, dtype=F32), F32(1e-6))
        tot += float(np.mean(np.abs(a - b), dtype=np.float64))
    return tot
