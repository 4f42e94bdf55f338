#!/usr/bin/env python
"""What the box's PCIe links + host memory deliver when N ranks copy at the same time: bare pinned-host H2D + D2H of
the e2e pipeline's byte counts (c2: 389 MB in, 87 MB out per step; c3: 347 MB each way), no kernels.

    python profiles/h2d_ceiling.py                                        (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/h2d_ceiling.py

Rank 0 prints one JSON line: per-rank and aggregate GB/s per direction, for copies issued alone (H2D only, D2H only)
and together (full duplex).  `bench.py` reports the same thing per step as `e2e.copy_only_ms_per_step`; this script
is the stand-alone record (profiles/r02_h2d_ceiling_*.json)."""
import json
import os

import torch
import torch.distributed as dist

rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
res = {}
for name, n_in, n_out in (("c2", 388694016, 86704128), ("c3", 347406336, 346816512)):
    h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n_in, dtype=torch.uint8, device=dev)
    d_out = torch.empty(n_out, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(mode, reps=8):
        def once():
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        for _ in range(2):
            once()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream(dev)
        e0.record()
        s1.wait_stream(cur)
        s2.wait_stream(cur)
        for _ in range(reps):
            once()
        cur.wait_stream(s1)
        cur.wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        t = torch.tensor([ms], device=dev)
        if world > 1:
            lst = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(lst, t)
            ms = max(float(x.item()) for x in lst)
        return ms

    r = {}
    for mode in ("h2d", "d2h", "both"):
        ms = run(mode)
        by_in = n_in if mode != "d2h" else 0
        by_out = n_out if mode != "h2d" else 0
        r[mode] = {"ms_per_step": ms, "h2d_gbs_per_gpu": by_in / ms / 1e6, "d2h_gbs_per_gpu": by_out / ms / 1e6,
                   "aggregate_gbs": (by_in + by_out) * world / ms / 1e6}
    res[name] = r
    del h_in, h_out, d_in, d_out
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
if rank == 0:
    print(json.dumps({"n_gpus": world, "host_cpus": len(os.sched_getaffinity(0)), "copies": res}))
