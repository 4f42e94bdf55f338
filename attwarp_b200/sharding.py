"""Sharding of a batch of independent images across the GPUs of one box.

The warp path has no cross-image dependency: every rank processes its
own images and there is NO collective on the data path.  ``torch.distributed`` (NCCL over
NVLink on GPUs, gloo in the CPU tests) is used only after the work, to gather per-rank timings
and 64-bit checksums.
"""

from __future__ import annotations

from typing import List, Sequence

import torch


def contiguous_shard(n_items: int, rank: int, world: int) -> range:
    """Items [start, stop) of rank ``rank``: sizes differ by at most one, order preserved."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def lpt_shard(costs: Sequence[float], world: int) -> List[List[int]]:
    """Greedy longest-processing-time assignment for ragged batches (BASELINE configs[3]):
    items sorted by decreasing cost (pixels in + pixels out), each given to the currently
    lightest rank.  Deterministic (ties broken by index / rank id).  Returns item indices per
    rank, each list in increasing index order."""
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    loads = [0.0] * world
    shards: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += float(costs[i])
    return [sorted(s) for s in shards]


def checksum64(t: torch.Tensor) -> int:
    """Position-weighted (so order-DEPENDENT) 64-bit checksum: sum of the bytes' values times
    ``(position % 251) + 1``.  Cheap, runs on the tensor's device; two runs that produce the same bytes in
    the same order -- a sharded and an unsharded pass over the same image -- give the same value."""
    flat = t.reshape(-1).view(torch.uint8).to(torch.int64)
    w = (torch.arange(flat.numel(), device=flat.device, dtype=torch.int64) % 251) + 1
    return int((flat * w).sum().item())


def gather_stats(elapsed_ms: float, n_images: int, checksum: int, device=None):
    """all_gather of (elapsed_ms, n_images, checksum) over the default process group.
    Returns a list of tuples, one per rank (a single tuple list when not distributed)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [(float(elapsed_ms), int(n_images), int(checksum))]
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" \
            else torch.device("cpu")
    mine = torch.tensor([int(round(float(elapsed_ms) * 1e6)), int(n_images), int(checksum)],
                        dtype=torch.int64, device=device)          # elapsed in nanoseconds
    out = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    res = []
    for t in out:
        v = t.cpu().tolist()
        res.append((v[0] / 1e6, int(v[1]), int(v[2])))
    return res


def aggregate_throughput(stats) -> float:
    """Whole-job images/s: all images of all ranks / the slowest rank's time."""
    total = sum(s[1] for s in stats)
    worst_ms = max(s[0] for s in stats)
    return total / (worst_ms / 1e3)
