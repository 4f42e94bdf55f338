"""CPU baseline runner: the reference's own CPU path, timed on host cores.

TEST / BENCH INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Used by ``bench.py`` for the
``cpu_baseline`` object and the ``--impl reference`` arm.  Two implementations of ONE composition:

* kind ``"reference"`` -- the UNMODIFIED reference functions, imported by ``oracle/ref_loader.py`` from
  ``/root/reference`` (build container) or from the byte-for-byte copy ``oracle/make_ref.py`` placed under the
  git-ignored ``baseline/_ref/`` (GPU box): ``MaskHookLogger._process_attention`` once per synthetic
  layer/step + ``finalize`` (llava.py:94-132) on the attention upcast to float32, then
  ``set_transform_function("identity")`` + ``warp_image_by_attention`` (new_method.py:198-283, 378-403);
* kind ``"port"`` -- the oracle restatement composed the same way (float32 hook expression, float64 NumPy
  marginals / cumsum, ``np.interp``-equivalent inversion, the real ``cv2.remap``), used when no copy of the
  reference is present.

Between the two sits one synthetic step that is not a reference function (BASELINE configs[1]/[2] "upsample
the token map to image resolution"): ``numpy_path.upsample_tokens_nearest``.

The reference is single-threaded Python per image; to give the CPU "all the host threads it can use" images
are spread over a fork()ed process pool (one image at a time per worker, ``cv2.setNumThreads(1)``,
``torch.set_num_threads(1)``).  ``threading_rows`` also times one process with the libraries' default
threading and with one thread.
"""

from __future__ import annotations

import contextlib
import io
import multiprocessing as mp
import os
import time

import numpy as np

from . import aggregate as OA
from . import numpy_path as ON
from . import ref_loader as RL

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


def reference_available() -> bool:
    return RL.available()


def _reference_modules():
    """(new_method, llava hooks) of the unmodified reference, imported once per process."""
    if "ref" not in _G:
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            _G["ref"] = (RL.new_method(), RL.llava_hooks())
    return _G["ref"]


def one_image_reference(i):
    import torch
    attn, tok, imgs, grid, out_hw, transform = (_G[k] for k in
                                                ("attn", "tok", "imgs", "grid", "out_hw", "transform"))
    nm, lh = _reference_modules()
    if attn is not None:
        L, Hh, T = attn.shape[1:]
        logger = lh.MaskHookLogger(None, "cpu")
        logger.set_image_token_range(0, T)
        a = torch.from_numpy(attn[i])                       # [L, Hh, T] float32: layers play the hooked steps
        for l in range(L):
            logger._process_attention(a[l].reshape(1, Hh, 1, T))
        t = logger.finalize().reshape(grid, grid).numpy()                          # stage 1
    else:
        t = tok[i]
    H, W = imgs.shape[1:3]
    full = ON.upsample_tokens_nearest(t, H, W)                                     # stage 2a (synthetic step)
    nm.set_transform_function(transform)
    out = nm.warp_image_by_attention(imgs[i], full, out_hw[1], out_hw[0])          # stages 2b-5
    return int(out[::7, ::7].sum())


def one_image_port(i):
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


def one_image(i):
    return one_image_reference(i) if _G.get("kind") == "reference" else one_image_port(i)


def _set_threads(n):
    """n = None leaves the libraries' defaults."""
    if n is None:
        return
    try:
        import cv2
        cv2.setNumThreads(n)
    except Exception:
        pass
    if _G.get("kind") == "reference":
        try:
            import torch
            torch.set_num_threads(max(n, 1))
        except Exception:
            pass


def _work(idx_range):
    _set_threads(1)
    return sum(one_image(i) for i in idx_range)


class CpuRunner:
    """Process pool over a fixed sample; ``step()`` processes the whole sample once.
    ``kind``: "reference" (default when a copy of the reference is present) or "port"."""

    def __init__(self, attn, tok, imgs, grid, out_hw, transform="identity", workers=None, kind=None):
        if kind is None:
            kind = "reference" if reference_available() else "port"
        if kind == "reference":
            _reference_modules()                      # import before the fork: the workers inherit the modules
        _G.update(attn=attn, tok=tok, imgs=imgs, grid=grid, out_hw=out_hw, transform=transform, kind=kind)
        self.kind = kind
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

    def threading_rows(self, n_images=4, repeats=3):
        """One process over the first ``n_images`` images: the libraries' default threading and one thread
        (median of ``repeats``).  Call BEFORE other steps of a single-worker runner or on any runner: it runs in
        the calling process."""
        import cv2
        import torch
        rows = {}
        n = min(n_images, self.n)
        default_cv2, default_torch = cv2.getNumThreads(), torch.get_num_threads()
        for name, threads in (("default_threading", None), ("single_thread", 1)):
            if threads is None:
                cv2.setNumThreads(default_cv2)
                torch.set_num_threads(default_torch)
            else:
                _set_threads(threads)
            ts = []
            for _ in range(repeats + 1):
                t0 = time.perf_counter()
                for i in range(n):
                    one_image(i)
                ts.append((time.perf_counter() - t0) / n)
            ts = sorted(ts[1:])
            rows[name] = {"images_per_s": 1.0 / ts[len(ts) // 2], "ms_per_image": ts[len(ts) // 2] * 1e3,
                          "cv2_threads": default_cv2 if threads is None else threads,
                          "torch_threads": default_torch if threads is None else threads}
        cv2.setNumThreads(default_cv2)
        torch.set_num_threads(default_torch)
        return rows

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()
            self.pool = None
