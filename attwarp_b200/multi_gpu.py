"""Multi-GPU entry points: one process per GPU, images sharded by index, no collective on the data path.

The warp has no cross-image dependency, so a box of 8 B200s runs 8 independent shards.
These helpers are what a driver such as ``AGW/main_batched.py:243-287`` (one ``warp_image_by_attention`` per image)
calls instead of its per-image loop when it runs under ``torchrun``:

    plan = multi_gpu.plan_ragged(sizes, out_sizes)              # same arguments on every rank -> same plan
    idx, outs = multi_gpu.warp_ragged_sharded(plan, tok_all, load_image)
    sums = multi_gpu.gather_image_checksums(plan, idx, outs)     # [n_images] int64, identical on every rank

``torch.distributed`` (NCCL over NVLink on the GPUs, gloo in the CPU tests) is used only AFTER the work, for
bookkeeping: per-image checksums (8 bytes per image) and per-rank timings.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import torch

from . import sharding


def _rank_world(rank, world):
    import torch.distributed as dist
    if rank is None or world is None:
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
        return 0, 1
    return int(rank), int(world)


@dataclass
class RaggedPlan:
    sizes: List[Tuple[int, int]]            # (H, W) per image
    out_sizes: List[Tuple[int, int]]        # (Ho, Wo) per image
    shards: List[List[int]]                 # image indices per rank, increasing
    world: int

    def mine(self, rank: Optional[int] = None) -> List[int]:
        r, _ = _rank_world(rank, self.world)
        return self.shards[r]

    def load(self) -> List[float]:
        """Cost (pixels in + pixels out) per rank: what greedy LPT balanced."""
        return [float(sum(self.sizes[i][0] * self.sizes[i][1] + self.out_sizes[i][0] * self.out_sizes[i][1]
                          for i in s)) for s in self.shards]


def plan_ragged(sizes: Sequence[Tuple[int, int]], out_sizes: Optional[Sequence[Tuple[int, int]]] = None,
                world: Optional[int] = None) -> RaggedPlan:
    """Greedy longest-processing-time split of a mixed-resolution batch (BASELINE configs[3]) by
    ``H*W + Ho*Wo``.  Deterministic: every rank computes the same plan from the same arguments."""
    _, w = _rank_world(0 if world is not None else None, world)
    sizes = [(int(h), int(w_)) for h, w_ in sizes]
    out_sizes = list(sizes) if out_sizes is None else [(int(h), int(w_)) for h, w_ in out_sizes]
    if len(out_sizes) != len(sizes):
        raise ValueError("plan_ragged: one output size per image")
    costs = [float(h * w_ + ho * wo) for (h, w_), (ho, wo) in zip(sizes, out_sizes)]
    return RaggedPlan(sizes, out_sizes, sharding.lpt_shard(costs, w), w)


def warp_ragged_sharded(plan: RaggedPlan, tok_all: torch.Tensor, load_image: Callable[[int], torch.Tensor],
                        rank: Optional[int] = None, device=None, transform="identity", exp_scale=1.0,
                        exp_divisor=1.0, apply_inverse=False, images=None, outs=None):
    """This rank's share of a ragged batch, stages 2-5, one launch per stage.

    tok_all      [n_images, gh, gw] float32 token maps (host or device; only this rank's rows are used)
    load_image   ``i -> uint8 HWC tensor`` of image i (host tensors are moved to the device); called for this
                 rank's images only -- pass ``images`` (already resident device tensors, in plan.mine() order)
                 to skip it
    Returns (indices, outs): this rank's image indices and their warped images (device tensors)."""
    from . import ops
    r, _ = _rank_world(rank, plan.world)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    idx = plan.shards[r]
    if not idx:
        return idx, []
    if images is None:
        images = [load_image(i).to(dev, non_blocking=True) for i in idx]
    for i, im in zip(idx, images):
        if tuple(im.shape[:2]) != plan.sizes[i]:
            raise ValueError(f"image {i} has shape {tuple(im.shape)} but the plan says {plan.sizes[i]}")
    tok = tok_all[torch.as_tensor(idx, device=tok_all.device)].to(dev)
    outs = ops.warp_ragged_from_tokens(tok, images, [plan.out_sizes[i] for i in idx], transform=transform,
                                       exp_scale=exp_scale, exp_divisor=exp_divisor,
                                       apply_inverse=apply_inverse, outs=outs)
    return idx, outs


def gather_image_checksums(plan: RaggedPlan, idx: Sequence[int], outs: Sequence[torch.Tensor],
                           device=None) -> torch.Tensor:
    """[n_images] int64 position-weighted checksums of every warped image of the batch, identical on every rank
    (each rank fills its own entries, one SUM all_reduce of 8 bytes per image merges them)."""
    import torch.distributed as dist
    n = len(plan.sizes)
    dev = outs[0].device if len(outs) else (torch.device(device) if device is not None else torch.device("cpu"))
    full = torch.zeros(n, dtype=torch.int64, device=dev)
    for i, o in zip(idx, outs):
        full[i] = sharding.checksum64(o)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() == "gloo":
            full = full.cpu()
        dist.all_reduce(full, op=dist.ReduceOp.SUM)
    return full


def warp_batch_sharded(attn: torch.Tensor, images: torch.Tensor, grid_hw, out_size=None, rank=None, world=None,
                       device=None, **kw):
    """Uniform batches (BASELINE configs[1]/[2]): this rank's contiguous slice of ``attn`` [B,L,Hh,T] and
    ``images`` [B,H,W,C] (host or device tensors holding the WHOLE batch) through the fused stages 1-5.
    Returns (range of image indices, warped images of that range on this rank's device)."""
    from . import ops
    r, w = _rank_world(rank, world)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    rng = sharding.contiguous_shard(images.shape[0], r, w)
    if len(rng) == 0:
        return rng, images[:0].to(dev)
    a = attn[rng.start:rng.stop].to(dev, non_blocking=True)
    im = images[rng.start:rng.stop].to(dev, non_blocking=True)
    return rng, ops.warp_from_attention_tokens(a, im, grid_hw, out_size, **kw)
