"""Batch pipelines: ``HostBatchPipeline`` is the call a user with NumPy/pinned-host data makes;
``StreamRing`` spreads consecutive device-resident batches over a few CUDA streams.

``HostBatchPipeline.run(attn_host, images_host, out_host)`` takes a whole batch living in
(pinned) host memory, splits it into chunks and drives, per chunk and on alternating CUDA
streams:  H2D(attention) -> H2D(images) -> fused stages 1-5 -> D2H(warped images), so the copy
engines and the SMs overlap.  Device buffers are allocated once and reused; nothing runs on the
CPU except the enqueueing.  This is what ``bench.py`` reports as ``e2e``.
"""

from __future__ import annotations

import torch

from . import ops


class StreamRing:
    """Round-robin over ``n`` CUDA streams for consecutive, independent batches that already live in HBM.

    The stages of one batch depend on each other, so on one stream they run in turn: stage 1
    (attention aggregation) saturates HBM while the SMs' issue slots idle, stage 5 (resample) is
    issue-bound while HBM idles.  Batches are independent (images never read each other's data,
    SURVEY.md section 8e), so batch k+1's stage 1 can share the GPU with batch k's stage 5: on a B200,
    BASELINE configs[1] goes from 120 us to ~91 us per batch with three or four streams.
    Every stream has its own workspace (``ops._workspace`` is keyed by stream); the caller must
    give concurrent batches distinct input/aux/output tensors.

        ring = StreamRing(4)
        ring.fork()                      # ring streams wait for the caller's stream
        for batch in batches:
            ring.submit(graph.replay)    # or any callable that enqueues on the current stream
        ring.join()                      # caller's stream waits for every ring stream
    """

    def __init__(self, n: int = 4, device=None, sm_share: int = 1):
        """``sm_share=2``: between ``fork()`` and ``join()`` every launch of the two streaming kernels fills only
        half of each SM (``attwarp_set_sm_share``), so that stage 1 of one batch and stage 5 of another are
        co-resident on every SM instead of overlapping only at their ramps (83 vs 90 us per batch at
        configs[1]; a single batch alone is slower that way, which is why it is not the default)."""
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.streams = [torch.cuda.Stream(self.device) for _ in range(max(1, int(n)))]
        self.k = 0
        self.sm_share = int(sm_share)
        self._prev_share = None

    def fork(self):
        caller = torch.cuda.current_stream(self.device)
        for s in self.streams:
            s.wait_stream(caller)
        if self.sm_share > 1:
            from ._lib import load
            self._prev_share = load().attwarp_set_sm_share(self.sm_share)

    def submit(self, fn):
        st = self.streams[self.k % len(self.streams)]
        self.k += 1
        with torch.cuda.stream(st):
            return fn()

    def join(self):
        caller = torch.cuda.current_stream(self.device)
        for s in self.streams:
            caller.wait_stream(s)
        if self._prev_share is not None:
            from ._lib import load
            load().attwarp_set_sm_share(self._prev_share)
            self._prev_share = None


class HostBatchPipeline:
    def __init__(self, chunk: int, L: int, Hh: int, grid_hw, image_hwc, out_hw=None,
                 attn_dtype=torch.bfloat16, img_dtype=torch.uint8, transform="identity",
                 n_streams: int = 2, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.chunk, self.L, self.Hh = chunk, L, Hh
        self.gh, self.gw = grid_hw
        self.H, self.W, self.C = image_hwc
        self.Ho, self.Wo = (self.H, self.W) if out_hw is None else out_hw
        self.transform = transform
        T = self.gh * self.gw
        self.streams = [torch.cuda.Stream(self.device) for _ in range(n_streams)]
        self.slots = []
        for _ in range(n_streams):
            self.slots.append(dict(
                attn=torch.empty(chunk, L, Hh, T, dtype=attn_dtype, device=self.device),
                img=torch.empty(chunk, self.H, self.W, self.C, dtype=img_dtype, device=self.device),
                out=torch.empty(chunk, self.Ho, self.Wo, self.C, dtype=img_dtype, device=self.device),
                aux=(torch.empty(chunk, T, dtype=torch.float32, device=self.device),
                     torch.empty(chunk, self.Wo, dtype=torch.float32, device=self.device),
                     torch.empty(chunk, self.Ho, dtype=torch.float32, device=self.device))))
        self.kernel_launches = 0

    def run(self, attn_host: torch.Tensor, images_host: torch.Tensor, out_host: torch.Tensor,
            tok_host: torch.Tensor | None = None):
        """attn_host [B,L,Hh,T], images_host [B,H,W,C], out_host [B,Ho,Wo,C]: host tensors
        (pinned for real overlap); ``tok_host`` [B,T] float32 optionally receives the stage-1 token maps (the
        attention map the reference drivers save next to the warped image, main.py:371).  Returns after
        everything is enqueued; call ``sync()``."""
        B = attn_host.shape[0]
        T = self.gh * self.gw
        if (tuple(attn_host.shape) != (B, self.L, self.Hh, T) or tuple(images_host.shape) != (B, self.H, self.W, self.C)
                or tuple(out_host.shape) != (B, self.Ho, self.Wo, self.C)
                or (tok_host is not None and (tuple(tok_host.shape) != (B, T) or tok_host.dtype != torch.float32))):
            raise ValueError("HostBatchPipeline.run: tensor shapes do not match the pipeline's configuration")
        if attn_host.dtype != self.slots[0]["attn"].dtype or images_host.dtype != self.slots[0]["img"].dtype \
                or out_host.dtype != self.slots[0]["out"].dtype:
            raise ValueError("HostBatchPipeline.run: tensor dtypes do not match the pipeline's configuration")
        caller = torch.cuda.current_stream(self.device)
        for s in self.streams:
            s.wait_stream(caller)
        k = 0
        for lo in range(0, B, self.chunk):
            hi = min(lo + self.chunk, B)
            n = hi - lo
            slot, st = self.slots[k % len(self.slots)], self.streams[k % len(self.streams)]
            with torch.cuda.stream(st):
                slot["attn"][:n].copy_(attn_host[lo:hi], non_blocking=True)
                slot["img"][:n].copy_(images_host[lo:hi], non_blocking=True)
                ops.warp_from_attention_tokens(
                    slot["attn"][:n], slot["img"][:n], (self.gh, self.gw), (self.Ho, self.Wo), "hwc",
                    transform=self.transform, out=slot["out"][:n],
                    aux=tuple(a[:n] for a in slot["aux"]))
                out_host[lo:hi].copy_(slot["out"][:n], non_blocking=True)
                if tok_host is not None:
                    tok_host[lo:hi].copy_(slot["aux"][0][:n], non_blocking=True)
            self.kernel_launches += 3
            k += 1
        for s in self.streams:
            caller.wait_stream(s)

    def sync(self):
        torch.cuda.current_stream(self.device).synchronize()


class HostTokenPipeline:
    """The same host-buffer pipeline for batches whose token maps already exist (BASELINE configs[2]: a [B,gh,gw]
    float32 token map per image instead of the attention tensor):  per chunk, on alternating streams,
    H2D(token maps) -> H2D(images) -> stages 2-5 (``maps_from_tokens`` + ``remap_bilinear``) -> D2H(warped
    images).  The images travel once in each direction, so the two PCIe directions work at the same time."""

    def __init__(self, chunk: int, grid_hw, image_hwc, out_hw=None, img_dtype=torch.uint8, transform="identity",
                 n_streams: int = 2, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.chunk = chunk
        self.gh, self.gw = grid_hw
        self.H, self.W, self.C = image_hwc
        self.Ho, self.Wo = (self.H, self.W) if out_hw is None else out_hw
        self.transform = transform
        self.streams = [torch.cuda.Stream(self.device) for _ in range(n_streams)]
        self.slots = []
        for _ in range(n_streams):
            self.slots.append(dict(
                tok=torch.empty(chunk, self.gh, self.gw, dtype=torch.float32, device=self.device),
                img=torch.empty(chunk, self.H, self.W, self.C, dtype=img_dtype, device=self.device),
                out=torch.empty(chunk, self.Ho, self.Wo, self.C, dtype=img_dtype, device=self.device),
                maps=(torch.empty(chunk, self.Wo, dtype=torch.float32, device=self.device),
                      torch.empty(chunk, self.Ho, dtype=torch.float32, device=self.device))))
        self.kernel_launches = 0

    def run(self, tok_host: torch.Tensor, images_host: torch.Tensor, out_host: torch.Tensor):
        """tok_host [B,gh,gw] float32, images_host [B,H,W,C], out_host [B,Ho,Wo,C] (pinned host tensors).
        Returns after everything is enqueued; call ``sync()``."""
        B = tok_host.shape[0]
        if (tuple(tok_host.shape) != (B, self.gh, self.gw) or tok_host.dtype != torch.float32
                or tuple(images_host.shape) != (B, self.H, self.W, self.C)
                or tuple(out_host.shape) != (B, self.Ho, self.Wo, self.C)
                or images_host.dtype != self.slots[0]["img"].dtype or out_host.dtype != self.slots[0]["out"].dtype):
            raise ValueError("HostTokenPipeline.run: tensors do not match the pipeline's configuration")
        caller = torch.cuda.current_stream(self.device)
        for s in self.streams:
            s.wait_stream(caller)
        k = 0
        for lo in range(0, B, self.chunk):
            hi = min(lo + self.chunk, B)
            n = hi - lo
            slot, st = self.slots[k % len(self.slots)], self.streams[k % len(self.streams)]
            with torch.cuda.stream(st):
                slot["tok"][:n].copy_(tok_host[lo:hi], non_blocking=True)
                slot["img"][:n].copy_(images_host[lo:hi], non_blocking=True)
                mx, my = slot["maps"][0][:n], slot["maps"][1][:n]
                ops.maps_from_tokens(slot["tok"][:n], (self.H, self.W), (self.Ho, self.Wo), self.transform,
                                     out=(mx, my))
                ops.remap_bilinear(slot["img"][:n], mx, my, "hwc", out=slot["out"][:n])
                out_host[lo:hi].copy_(slot["out"][:n], non_blocking=True)
            self.kernel_launches += 2
            k += 1
        for s in self.streams:
            caller.wait_stream(s)

    def sync(self):
        torch.cuda.current_stream(self.device).synchronize()


class HostCopyProbe:
    """The copies of a host pipeline WITHOUT its kernels: the same chunks, byte counts, streams and pinned
    buffers, so that ``bench.py`` can report how far ``e2e`` is from what the box's PCIe / host memory delivers
    when every rank copies at the same time (``e2e.copy_only_ms_per_step``)."""

    def __init__(self, pipeline):
        self.p = pipeline

    def run(self, inputs_host, out_host):
        """inputs_host: the host tensors ``run`` copies in, in order, matched to the slot buffers by name."""
        p = self.p
        names = ("attn", "img") if "attn" in p.slots[0] else ("tok", "img")
        B = inputs_host[0].shape[0]
        caller = torch.cuda.current_stream(p.device)
        for s in p.streams:
            s.wait_stream(caller)
        k = 0
        for lo in range(0, B, p.chunk):
            hi = min(lo + p.chunk, B)
            n = hi - lo
            slot, st = p.slots[k % len(p.slots)], p.streams[k % len(p.streams)]
            with torch.cuda.stream(st):
                for nm, h in zip(names, inputs_host):
                    slot[nm][:n].copy_(h[lo:hi], non_blocking=True)
                out_host[lo:hi].copy_(slot["out"][:n], non_blocking=True)
            k += 1
        for s in p.streams:
            caller.wait_stream(s)
