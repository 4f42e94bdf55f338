"""CPU: host-side sharding logic, incl. a world_size-2 gloo run of the stats gather that the
multi-GPU bench uses over NCCL."""

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from attwarp_b200 import sharding


def test_contiguous_shard_covers_everything():
    for n in (0, 1, 7, 256, 1024):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                seen += list(sharding.contiguous_shard(n, r, world))
            assert seen == list(range(n))
            sizes = [len(sharding.contiguous_shard(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_lpt_shard_balances_mixed_resolutions():
    rng = np.random.default_rng(1237)
    sides = rng.integers(224, 2049, 1024)
    costs = (sides.astype(np.int64) ** 2 * 2).tolist()
    shards = sharding.lpt_shard(costs, 8)
    assert sorted(i for s in shards for i in s) == list(range(1024))
    loads = [sum(costs[i] for i in s) for s in shards]
    assert max(loads) / (sum(loads) / 8) < 1.01          # within 1 % of perfect balance
    assert shards == sharding.lpt_shard(costs, 8)         # deterministic


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        idx = sharding.contiguous_shard(10, rank, world)
        data = torch.arange(1000, dtype=torch.uint8).reshape(10, 100)[idx.start:idx.stop]
        stats = sharding.gather_stats(5.0 + rank, len(idx), sharding.checksum64(data))
        q.put((rank, stats, sharding.aggregate_throughput(stats)))
    finally:
        dist.destroy_process_group()


def test_gather_stats_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1]                          # every rank sees the same table
    stats = res[0][1]
    assert [s[1] for s in stats] == [5, 5] and abs(stats[1][0] - 6.0) < 1e-6
    assert abs(res[0][2] - 10 / 6e-3) < 1e-6                # all images / slowest rank
    full = torch.arange(1000, dtype=torch.uint8).reshape(10, 100)
    assert stats[0][2] == sharding.checksum64(full[:5]) and stats[1][2] == sharding.checksum64(full[5:])


def _worker_plan(rank, world, port, q):
    """multi_gpu bookkeeping over gloo: every rank derives the same LPT plan, fills in the checksums of ITS images
    (CPU stand-ins for the warped images), one SUM all_reduce merges them."""
    from attwarp_b200 import multi_gpu
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        sizes = [(int(h), int(w)) for h, w in rng.integers(8, 64, (11, 2))]
        plan = multi_gpu.plan_ragged(sizes)
        idx = plan.mine()
        outs = [torch.full((h, w, 3), i % 251, dtype=torch.uint8) for i, (h, w) in ((i, sizes[i]) for i in idx)]
        sums = multi_gpu.gather_image_checksums(plan, idx, outs)
        q.put((rank, plan.shards, sums.tolist()))
    finally:
        dist.destroy_process_group()


def test_multi_gpu_plan_and_checksums_world2_gloo():
    from attwarp_b200 import multi_gpu
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_plan, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1] and res[0][2] == res[1][2]         # same plan, same merged table on both ranks
    shards = res[0][1]
    assert sorted(i for s in shards for i in s) == list(range(11))
    rng = np.random.default_rng(7)
    sizes = [(int(h), int(w)) for h, w in rng.integers(8, 64, (11, 2))]
    want = [sharding.checksum64(torch.full((h, w, 3), i % 251, dtype=torch.uint8)) for i, (h, w) in enumerate(sizes)]
    assert res[0][2] == want                                         # == the unsharded table
    plan1 = multi_gpu.plan_ragged(sizes, world=1)
    assert plan1.shards == [list(range(11))]
    loads = multi_gpu.plan_ragged(sizes, world=2).load()
    assert max(loads) / (sum(loads) / 2) < 1.15
