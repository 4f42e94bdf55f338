"""Oracle: restatement of the hook-logger attention reducers (stage 1).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Follows ``/root/reference/Attention Guided Warping/attention_extraction/llava.py``:

* ``MaskHookLogger._process_attention`` ...... ``llava.py:94-116``
* ``MaskHookLogger.finalize`` ................ ``llava.py:124-132``
* ``BatchMaskHookLogger._process_attention`` . ``llava.py:385-396``
* ``BatchMaskHookLogger.finalize_batch`` ..... ``llava.py:401-411``
* mask post-processing (``normalize``/``enhance``/``revise_mask``) ``llava.py:207-238``

Parity is defined on the input upcast to float32: the reference divides
and averages in the tensor's own dtype; fp16/bf16 evaluation deviates by 5e-4 / 3e-3 relative,
far above the 1e-5 tolerance, so the oracle (and the CUDA kernel) upcast first.
"""

from __future__ import annotations

import numpy as np

F32 = np.float32
NUM_IMAGE_TOKENS = 576      # llava.py:50,351 (24x24 patches)


def renorm_head_mean(rows):
    """rows [..., Hh, T] float32 -> [..., T]: a / (sum_t a + 1e-12), mean over heads."""
    rows = np.asarray(rows, dtype=F32)
    s = rows.sum(axis=-1, keepdims=True, dtype=F32) + F32(1e-12)
    return (rows / s).astype(F32).mean(axis=-2, dtype=F32)


def aggregate_attention(attn, tok_start=None, T=None):
    """Closed form used by the synthetic benchmark (SURVEY.md 'fact 3'):

    ``out[b,t] = mean_l mean_h ( a[b,l,h,st_b+t] / (sum_t a[b,l,h,st_b+t] + 1e-12) )``

    attn: [B, L, Hh, K] (any float dtype, upcast to float32); tok_start: optional [B] ints;
    T: slice length (default K)."""
    a = np.asarray(attn).astype(F32)
    B, L, Hh, K = a.shape
    if T is None:
        T = K
    if tok_start is None:
        tok_start = np.zeros(B, dtype=np.int64)
    out = np.empty((B, T), dtype=F32)
    for b in range(B):
        st = int(tok_start[b])
        per_step = renorm_head_mean(a[b, :, :, st:st + T])      # [L, T]
        out[b] = per_step.mean(axis=0, dtype=F32)
    return out


class HookAggregator:
    """Drives like ``MaskHookLogger`` (B == 1) / ``BatchMaskHookLogger`` (any B)."""

    def __init__(self, num_image_tokens=NUM_IMAGE_TOKENS):
        self.num_image_tokens = num_image_tokens
        self.starts = None
        self.ends = None
        self.steps = []

    def set_ranges(self, starts, ends):
        assert len(starts) == len(ends)
        self.starts, self.ends = list(starts), list(ends)

    def process(self, attn_weights):
        """attn_weights [B, Hh, q, kv]: last query row, per-sample [st,ed) slice."""
        a = np.asarray(attn_weights).astype(F32)
        B, Hh, q, kv = a.shape
        per_sample = []
        for b in range(B):
            if self.starts is None:
                st, ed = 1, min(1 + self.num_image_tokens, kv)        # llava.py:99-102
            else:
                st, ed = self.starts[b], min(self.ends[b], kv)
            per_sample.append(renorm_head_mean(a[b, :, -1, st:ed]))
        self.steps.append(np.stack(per_sample, axis=0))

    def finalize(self):
        """[B, T] step mean; uniform 1/T when nothing was captured."""
        if not self.steps:
            B = 1 if self.starts is None else len(self.starts)
            return np.full((B, self.num_image_tokens), 1.0 / self.num_image_tokens, dtype=F32)
        return np.stack(self.steps, axis=0).mean(axis=0, dtype=F32)


# ---- mask post-processing (llava.py:207-238) -- 'next' row N2 -------------------------
def revise_mask(patch_mask, kernel_size=3, enhance_coe=10):
    """[24,24] float32 -> [24,24]: min-max, z-score*coe, sigmoid, k x k box (replicate pad)."""
    m = np.asarray(patch_mask, dtype=F32)
    m = (m - m.min()) / (m.max() - m.min())
    m = m - m.mean(dtype=F32)
    m = m / m.std(ddof=1, dtype=F32)          # torch.std is unbiased
    m = (m * F32(enhance_coe)).astype(F32)
    m = (F32(1) / (F32(1) + np.exp(-m))).astype(F32)
    m = np.clip(m, 0, 1)
    pad = (kernel_size - 1) // 2
    mp = np.pad(m, pad, mode="edge")
    out = np.zeros_like(m)
    w = F32(1.0 / kernel_size ** 2)
    for dy in range(kernel_size):
        for dx in range(kernel_size):
            out += mp[dy:dy + m.shape[0], dx:dx + m.shape[1]] * w
    return out.astype(F32)
