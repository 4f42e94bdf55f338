#!/usr/bin/env python
"""Host + device cost of one hooked attention step (MaskHookLogger / BatchMaskHookLogger._process_attention) on a live
[B, Hh, q, kv] tensor, with kv growing from step to step like under generate() with a KV cache.  Wall clock per call
over many calls (the kernel is asynchronous: this is what the hook adds to the model's own launch stream), and the same
with ATTWARP_HOOK_FAST=0 (every step through ops.aggregate_attention).  Never a bench number.

    python profiles/hook_overhead.py
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from attwarp_b200 import attention_extraction as AE  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)


class _M:           # the loggers only keep a reference to the model
    pass


def run(cls, B, n=400):
    big = torch.rand(B, 32, 1, 700 + n, device=dev, generator=g).half()
    lg = cls(_M(), dev)
    if cls is AE.BatchMaskHookLogger:
        lg.set_batch_image_token_ranges([35] * B, [35 + 576] * B)
    else:
        lg.set_image_token_range(35, 35 + 576)
    views = [big[:, :, :, : 700 + i] for i in range(n)]          # kv grows by one per step (non-contiguous heads)
    for v in views[:20]:
        lg._process_attention(v)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for v in views[20:]:
        lg._process_attention(v)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return (t1 - t0) / (n - 20) * 1e6, (t2 - t0) / (n - 20) * 1e6


for cls, B in ((AE.MaskHookLogger, 1), (AE.BatchMaskHookLogger, 16)):
    enq, tot = run(cls, B)
    print(f"{cls.__name__:22s} B={B:2d} [B,32,1,kv>=700] fp16: {enq:6.1f} us per hooked step to enqueue, {tot:6.1f} us with the device drained"
          f"  (ATTWARP_HOOK_FAST={os.environ.get('ATTWARP_HOOK_FAST', '1')})", flush=True)
