"""CPU baseline runner: the oracle port of the reference path, timed on host cores.

TEST / BENCH INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Used by ``bench.py`` for the
``cpu_baseline`` object and the ``--impl reference`` arm.  The reference itself is pure Python
over NumPy/OpenCV and cannot travel to the GPU box (``/root/reference`` does not exist there),
so what is timed is the oracle restatement composed exactly like the reference composes its
library calls: float32 hook-logger expression (llava.py:109-114,131-132), then
``warp_image_by_attention`` (new_method.py:198-283) = float64 NumPy marginals/cumsum,
``np.interp``-equivalent inversion, ``np.meshgrid`` + the real ``cv2.remap`` -- kind "port".
The reference is single-threaded per image; to give the CPU "all the host threads it can use"
images are spread over a fork()ed process pool (one image at a time per worker,
``cv2.setNumThreads(1)``).
"""

from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np

from . import aggregate as OA
from . import numpy_path as ON

_G = {}


def make_c2_sample(n_images, L=32, Hh=32, grid=24, side=336, seed=1235):
    """Synthetic inputs of BASELINE configs[1] (same distributions as the GPU bench)."""
    rng = np.random.default_rng(seed)
    T = grid * grid
    logits = rng.standard_normal((n_images, L, Hh, T), dtype=np.float32) * 2
    logits -= logits.max(-1, keepdims=True)
    attn = np.exp(logits)
    attn /= attn.sum(-1, keepdims=True)
    imgs = rng.integers(0, 256, (n_images, side, side, 3), dtype=np.uint8)
    return attn, imgs


def make_c3_sample(n_images, grid=48, side=1344, seed=1236):
    rng = np.random.default_rng(seed)
    tok = rng.random((n_images, grid, grid)).astype(np.float32) ** 3
    tok /= tok.sum(axis=(1, 2), keepdims=True)
    imgs = rng.integers(0, 256, (n_images, side, side, 3), dtype=np.uint8)
    return tok, imgs


def one_image(i):
    attn, tok, imgs, grid, out_hw, transform = (_G[k] for k in
                                                ("attn", "tok", "imgs", "grid", "out_hw", "transform"))
    if attn is not None:
        t = OA.aggregate_attention(attn[i:i + 1])[0].reshape(grid, grid)          # stage 1
    else:
        t = tok[i]
    H, W = imgs.shape[1:3]
    full = ON.upsample_tokens_nearest(t, H, W)                                     # stage 2a
    out = ON.warp_image_by_attention(imgs[i], full, out_hw[1], out_hw[0], transform,
                                     remap_backend="cv2")                          # stages 2b-5
    return int(out[::7, ::7].sum())


def _work(idx_range):
    try:
        import cv2
        cv2.setNumThreads(1)
    except Exception:
        pass
    return sum(one_image(i) for i in idx_range)


class CpuRunner:
    """Process pool over a fixed sample; ``step()`` processes the whole sample once."""

    def __init__(self, attn, tok, imgs, grid, out_hw, transform="identity", workers=None):
        _G.update(attn=attn, tok=tok, imgs=imgs, grid=grid, out_hw=out_hw, transform=transform)
        self.n = imgs.shape[0]
        avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
        self.workers = max(1, min(workers or avail, self.n))
        self.pool = None
        if self.workers > 1:
            self.pool = mp.get_context("fork").Pool(self.workers)
        per = -(-self.n // (self.workers * 4))
        self.tasks = [range(lo, min(lo + per, self.n)) for lo in range(0, self.n, per)]

    def step(self):
        t0 = time.perf_counter()
        if self.pool is None:
            chk = sum(_work(t) for t in self.tasks)
        else:
            chk = sum(self.pool.map(_work, self.tasks))
        return time.perf_counter() - t0, chk

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()
            self.pool = None
