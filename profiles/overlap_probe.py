#!/usr/bin/env python
"""Probe: do consecutive c2 steps overlap when they alternate between two streams?
(stage 1 is HBM-bound, stage 5 issue-bound: co-running them should beat running them in turn)

    [ATTWARP_REMAP_CTAS_PER_SM=2] python profiles/overlap_probe.py [--streams 2 --steps 60]
"""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from attwarp_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=2)
ap.add_argument("--steps", type=int, default=60)
ap.add_argument("--sets", type=int, default=4)
ap.add_argument("--graph", type=int, default=0)
a = ap.parse_args()
dev = torch.device("cuda", 0)
B, L, Hh, G, S, C = 256, 32, 32, 24, 336, 3
gen = torch.Generator(device=dev).manual_seed(1)
sets = []
for _ in range(a.sets):
    z = [torch.softmax(torch.randn(64, L, Hh, G * G, device=dev, generator=gen) * 2, -1).to(torch.bfloat16) for _ in range(B // 64)]
    sets.append(dict(attn=torch.cat(z), img=torch.randint(0, 256, (B, S, S, C), device=dev, dtype=torch.uint8, generator=gen),
                     out=torch.empty(B, S, S, C, device=dev, dtype=torch.uint8),
                     aux=(torch.empty(B, G * G, device=dev), torch.empty(B, S, device=dev), torch.empty(B, S, device=dev))))
    del z
streams = [torch.cuda.Stream(device=dev) for _ in range(a.streams)]

def enqueue(s):
    ops.warp_from_attention_tokens(s["attn"], s["img"], (G, G), None, "hwc", transform="identity", out=s["out"], aux=s["aux"])

graphs = [ops.GraphedCall(lambda s=s: enqueue(s), device=dev) for s in sets] if a.graph else None

def run(n):
    for i in range(n):
        with torch.cuda.stream(streams[i % a.streams]):
            if graphs:
                graphs[i % a.sets].replay()
            else:
                enqueue(sets[i % a.sets])

run(8)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for st in streams:
    st.wait_event(e0)
run(a.steps)
for st in streams:
    torch.cuda.current_stream().wait_stream(st)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
print(f"graph={a.graph} sets={a.sets} streams={a.streams} remap_ctas_per_sm={os.environ.get('ATTWARP_REMAP_CTAS_PER_SM', 'max')}: {ms * 1e3:.1f} us/step, {B / ms * 1e3:.0f} img/s")
