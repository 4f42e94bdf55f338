#!/usr/bin/env python
"""bench.py -- warped images/s of the attention-guided warp hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic input:
  c2 (default; BASELINE.json configs[1]): 256 x 336^2 RGB uint8 + bf16 attention
      [256, 32, 32, 576]  -> stage 1 aggregation -> 24x24 token map -> marginals / CDF /
      inverse-CDF maps -> bilinear resample (cv2.remap semantics) to 336^2.
  c3 (configs[2]): 64 x 1344^2 RGB uint8 + 48x48 token maps -> maps -> resample.
  c4 (configs[3]): 1024 mixed-resolution images, LPT-sharded over the ranks (strong scaling), ragged launches.
  c5 (configs[4]): 128 x 3 x 512^2 float32 images + PDFs [128,24] x 2 -> fused PDF->CDF -> maps -> resample.
For N > 1 (launched under torchrun, one rank per GPU) every rank processes its own batch
(weak scaling, images sharded by index, no data-path collective); NCCL only gathers timings and
checksums.  Rank 0 prints ONE JSON line.

`value`  : whole-job images/s, inputs resident in HBM, CUDA-event timed, max over ranks.
`e2e`    : the same metric through the host-buffer pipeline (pinned host -> H2D -> kernels -> D2H).
`roofline`: the dominant kernel's algorithmic bytes / its CUDA-event duration vs the measured
            HBM copy peak (MEASURED_PEAKS.json).
`cpu_baseline`: the oracle port of the reference CPU path on this box's host cores (bounded sample).
`--impl reference`: only the CPU arm, same metric/config, `"impl": "reference"`.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "warped_images_per_sec"
UNIT = "images/s"

WORKLOADS = {
    "c2": dict(name="c2: LLaVA-1.5 batch 256x336^2 RGB u8 + bf16 attention [256,32,32,576] "
                    "aggregation + 24x24 token map -> inverse-CDF warp -> 336^2",
               B=256, L=32, Hh=32, grid=24, side=336, C=3, has_attention=True),
    "c3": dict(name="c3: Qwen-VL-style batch 64x1344^2 RGB u8 + 48x48 token map upsample + "
                    "inverse-CDF warp -> 1344^2",
               B=64, L=0, Hh=0, grid=48, side=1344, C=3, has_attention=False),
    "c4": dict(name="c4: mixed-resolution batch of 1024 RGB u8 images (sides uniform in [224, 2048]) + 24x24 "
                    "token maps -> inverse-CDF warp at input size, ragged launches, LPT-sharded over the GPUs",
               B=1024, L=0, Hh=0, grid=24, side=0, C=3, has_attention=False, ragged=True),
    "c5": dict(name="c5: MarginalNet-style PDFs (2 x [128,24], softmax of seeded logits, alpha=0.1) -> fused "
                    "PDF->CDF + inverse-CDF maps + float32 resample of 128x3x512^2 images (NCHW)",
               B=128, L=0, Hh=0, grid=24, side=512, C=3, has_attention=False, pdf=True),
}


class _StdoutToStderr:
    """OS-level redirect of fd 1 to fd 2 while libraries that printf() to stdout initialise (NCCL prints its
    version banner with a raw printf at NCCL_DEBUG=VERSION): stdout must carry the ONE JSON line only."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def init_nccl(dev):
    import torch.distributed as dist
    with _StdoutToStderr():
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()                       # communicator creation (and its banner) happens here at the latest


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy, burst)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


def ncu_traffic(kernel, workload):
    """Per-launch DRAM bytes of `kernel` from the committed ncu capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(workload, {}).get(kernel)
    except Exception:
        return None


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
def _gpu_uuid(index):
    try:
        import torch
        return str(torch.cuda.get_device_properties(index).uuid)
    except Exception:
        return None


class ClockSampler:
    """SM clock and clock-event (throttle) reasons DURING the timed region.  The region is a few
    milliseconds long, so the sampler is an NVML polling thread (a sample every ~0.2 ms; every NVML
    call releases the GIL); `nvidia-smi -lms` (50 ms period) is the fallback when NVML cannot be loaded."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, uuid=None):
        self.proc = None
        self.path = None
        self.gpu_index = gpu_index
        self.uuid = uuid
        self.thread = None
        self.samples = []
        self._stop = False
        self.nvml = None
        self.handle = None

    def _nvml_init(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if self.uuid:
                for cand in (self.uuid, "GPU-" + self.uuid):
                    try:
                        h = pynvml.nvmlDeviceGetHandleByUUID(cand.encode() if isinstance(cand, str) else cand)
                        break
                    except Exception:
                        h = None
            if h is None:
                idx = self.gpu_index
                vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
                try:
                    if vis:
                        idx = int(vis.split(",")[self.gpu_index])
                except Exception:
                    pass
                h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nvml, self.handle = pynvml, h
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            return True
        except Exception:
            self.nvml = None
            return False

    def _poll(self):
        nv, h = self.nvml, self.handle
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
        while not self._stop:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                bits = int(get_reasons(h)) if get_reasons else 0
                self.samples.append((mhz, bits))
            except Exception:
                pass
            time.sleep(0.0002)

    def start(self):
        if self._nvml_init():
            import threading
            # the main thread enqueues steps in a tight Python loop: with the default 5 ms GIL switch interval
            # the polling thread would get one sample per 5 ms
            self._switch = sys.getswitchinterval()
            sys.setswitchinterval(1e-4)
            self._stop = False
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-i", str(self.gpu_index), "-lms", "50"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            self._stop = True
            self.thread.join(timeout=2)
            sys.setswitchinterval(self._switch)
            nv = self.nvml
            table = {}
            for nm, attrs in (("hw_slowdown", ("nvmlClocksEventReasonHwSlowdown", "nvmlClocksThrottleReasonHwSlowdown")),
                              ("hw_thermal_slowdown", ("nvmlClocksEventReasonHwThermalSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown")),
                              ("sw_thermal_slowdown", ("nvmlClocksEventReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwThermalSlowdown")),
                              ("sw_power_cap", ("nvmlClocksEventReasonSwPowerCap", "nvmlClocksThrottleReasonSwPowerCap"))):
                for a in attrs:
                    if hasattr(nv, a):
                        table[nm] = int(getattr(nv, a))
                        break
            if not self.samples:
                return None
            sm = sorted(s[0] for s in self.samples)
            reasons = sorted(nm for nm, bit in table.items() if any(s[1] & bit for s in self.samples))
            return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons,
                    "samples": len(sm), "how": "NVML polled during the timed region"}
        if self.proc is None:
            return None
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for k, nm in enumerate(names):
                    if f[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if not sm:
            return None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "how": "nvidia-smi -lms 50 during the timed region"}


# ------------------------------------------------------------------------------------------
# CPU arm (oracle port of the reference)
# ------------------------------------------------------------------------------------------
def cpu_arm(wl_key, steps, warmup, budget_s, full_steps):
    """Times the oracle port on the host cores.  Each step = one pass over a bounded sample of
    the workload (the full batch when it is cheap enough)."""
    from oracle import cpu_baseline as CB
    wl = WORKLOADS[wl_key]
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    # calibrate on a couple of images, single process
    if wl["has_attention"]:
        attn, imgs = CB.make_c2_sample(2, wl["L"], wl["Hh"], wl["grid"], wl["side"])
        tok = None
    else:
        tok, imgs = CB.make_c3_sample(2, wl["grid"], wl["side"])
        attn = None
    r = CB.CpuRunner(attn, tok, imgs, wl["grid"], (wl["side"], wl["side"]), workers=1)
    r.step()
    t1, _ = r.step()
    per_img = t1 / 2
    r.close()
    total_steps = steps + warmup
    # images per step so that the whole run fits the budget (assume ~70 % parallel efficiency)
    n = int(budget_s / max(total_steps, 1) / per_img * avail * 0.7)
    n = max(min(n, wl["B"]), min(avail, wl["B"]), 1)
    if wl["has_attention"]:
        attn, imgs = CB.make_c2_sample(n, wl["L"], wl["Hh"], wl["grid"], wl["side"])
    else:
        tok, imgs = CB.make_c3_sample(n, wl["grid"], wl["side"])
    runner = CB.CpuRunner(attn, tok, imgs, wl["grid"], (wl["side"], wl["side"]))
    for _ in range(warmup):
        runner.step()
    t = 0.0
    for _ in range(steps):
        dt, _ = runner.step()
        t += dt
    runner.close()
    value = n * steps / t
    import cv2
    import numpy
    info = {"value": value, "unit": UNIT, "cores": runner.workers, "kind": "port",
            "sample": f"{n} of {wl['B']} images per step x {steps} steps ({warmup} warm-up), "
                      f"{runner.workers} fork()ed workers x 1 image at a time, cv2.setNumThreads(1); "
                      f"single-process cost {per_img * 1e3:.2f} ms/image; numpy {numpy.__version__}, "
                      f"cv2 {cv2.__version__}; host cpus visible {avail}"}
    return info, t / steps * 1e3


def run_reference(args, rank, world):
    if rank != 0:
        return 0
    wl = WORKLOADS[args.workload]
    if wl.get("ragged") or wl.get("pdf"):
        print(json.dumps({"impl": "reference", "unavailable": "the CPU arm covers the bench workloads c2 and c3; "
                          "c4 and c5 are parity/scaling configurations"}), flush=True)
        return 0
    info, ms_per_step = cpu_arm(args.workload, args.steps, args.warmup, budget_s=120.0,
                                full_steps=True)
    line = {"impl": "reference", "metric": METRIC, "value": info["value"], "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64+u8", "data": "synthetic",
            "config": {"workload": wl["name"], "timing": "wall clock around each CPU step"},
            "cpu_baseline": info,
            "e2e": {"value": info["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def run_gpu_ragged(args, rank, local_rank, world):
    """configs[3]: ONE batch of 1024 mixed-resolution images split over the ranks by greedy LPT on
    pixels in + pixels out (strong scaling, no data-path collective); a step = stages 2-5 over the
    rank's shard through the ragged entry point (one launch per stage)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from attwarp_b200 import ops, sharding

    wl = WORKLOADS[args.workload]
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        init_nccl(dev)
    B, C, grid = wl["B"], wl["C"], wl["grid"]
    sides = np.random.default_rng(1237).integers(224, 2049, size=B)
    shards = sharding.lpt_shard([2.0 * float(s) * float(s) for s in sides], world)
    mine = shards[rank]
    gen = torch.Generator(device=dev).manual_seed(1237 + rank)
    imgs = [torch.randint(0, 256, (int(sides[i]), int(sides[i]), C), device=dev, dtype=torch.uint8, generator=gen)
            for i in mine]
    outs = [torch.empty_like(t) for t in imgs]
    tok = torch.rand(len(mine), grid, grid, device=dev, generator=gen) ** 3
    tok = (tok / tok.sum(dim=(1, 2), keepdim=True)).contiguous()
    my_bytes = 2 * sum(t.numel() for t in imgs)

    def step():
        ops.warp_ragged_from_tokens(tok, imgs, outs=outs)
        return 2

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank, _gpu_uuid(local_rank))
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    ev0.record()
    for _ in range(args.steps):
        launches += step()
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = ev0.elapsed_time(ev1)
    chk = sharding.checksum64(outs[0])
    stats = sharding.gather_stats(elapsed_ms, len(mine) * args.steps, chk, dev)
    bstats = sharding.gather_stats(elapsed_ms, my_bytes // 1024, 0, dev)
    value = sharding.aggregate_throughput(stats)
    worst_ms = max(s[0] for s in stats)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0
    peak, peak_src = measured_peak()
    total_bytes = sum(b[1] for b in bstats) * 1024
    gbs = total_bytes / (worst_ms / args.steps) / 1e6 / world
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": worst_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64 (stages 2-4), u8 fixed-point (stage 5)",
            "data": "synthetic",
            "config": {"workload": wl["name"], "images_total": B,
                       "images_per_rank": [len(s) for s in shards],
                       "l2_policy": f"each rank's shard is {my_bytes / 1e6:.0f} MB in+out > 126 MB L2",
                       "parallelism": f"LPT shards over {world} GPU(s), no data-path collective",
                       "launch": "one maps + one resample launch per step (descriptor table built on the host each step)"},
            "clocks": clocks, "e2e": None, "gpu_launches": launches * world,
            "roofline": {"bound": "hbm", "kernel": "maps_from_tokens + remap_u8_stream (whole step, host table build included)",
                         "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": None,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": total_bytes // world},
            "cpu_baseline": None, "per_rank_ms": [s[0] for s in stats]}
    print(json.dumps(line), flush=True)
    return 0


def run_gpu_pdf(args, rank, local_rank, world):
    """configs[4]: predicted PDFs feeding the fused CDF + resample path (trainer.py:212-218, 285-289 chain):
    a step = attwarp_warp_from_pdfs over one batch of 128 float32 NCHW images (three launches)."""
    import torch
    import torch.distributed as dist

    from attwarp_b200 import ops, sharding
    from attwarp_b200.batched import StreamRing

    wl = WORKLOADS[args.workload]
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        init_nccl(dev)
    B, C, side, N = wl["B"], wl["C"], wl["side"], wl["grid"]
    R = args.rotate if args.rotate > 0 else 3
    gen = torch.Generator(device=dev).manual_seed(1238 + rank)
    sets = []
    for _ in range(R):
        sets.append(dict(px=torch.softmax(torch.randn(B, N, device=dev, generator=gen) * 1.5, -1),
                         py=torch.softmax(torch.randn(B, N, device=dev, generator=gen) * 1.5, -1),
                         img=torch.rand(B, C, side, side, device=dev, generator=gen),
                         out=torch.empty(B, C, side, side, device=dev)))

    def enqueue(s):
        ops.warp_from_pdfs(s["img"], s["px"], s["py"], alpha=0.1, layout="chw", out=s["out"])

    graphs = None if args.no_graph else [ops.GraphedCall(lambda s=s: enqueue(s), device=dev) for s in sets]
    n_streams = max(1, min(args.streams if args.streams > 0 else 3, R))
    ring = StreamRing(n_streams, dev) if n_streams > 1 else None

    def run_step(i):
        if graphs is not None:
            graphs[i % R].replay()
        else:
            enqueue(sets[i % R])

    def step(i):
        if ring is not None:
            ring.submit(lambda: run_step(i))
        else:
            run_step(i)
        return 3

    if ring is not None:
        ring.fork()
    for i in range(args.warmup):
        step(i)
    if ring is not None:
        ring.join()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank, _gpu_uuid(local_rank))
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    ev0.record()
    if ring is not None:
        ring.fork()
    for i in range(args.steps):
        launches += step(args.warmup + i)
    if ring is not None:
        ring.join()
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = ev0.elapsed_time(ev1)
    chk = int(sets[(args.warmup + args.steps - 1) % R]["out"][:2].double().sum().item() * 1e3)
    stats = sharding.gather_stats(elapsed_ms, B * args.steps, chk, dev)
    value = sharding.aggregate_throughput(stats)
    worst_ms = max(s[0] for s in stats)
    # the resample kernel alone: back-to-back launches with the maps of the last step
    _, _, _, mx, my = ops.warp_from_pdfs(sets[0]["img"], sets[0]["px"], sets[0]["py"], alpha=0.1, layout="chw",
                                         out=sets[0]["out"], return_aux=True)
    kreps = max(10, min(args.steps, 30))
    for i in range(3):
        ops.remap_bilinear(sets[i % R]["img"], mx, my, "chw", out=sets[i % R]["out"])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(kreps):
        ops.remap_bilinear(sets[i % R]["img"], mx, my, "chw", out=sets[i % R]["out"])
    e1.record()
    torch.cuda.synchronize()
    kms = e0.elapsed_time(e1) / kreps
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0
    peak, peak_src = measured_peak()
    by = B * C * side * side * 4 * 2
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": worst_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 (PDF/CDF rows), f64 (inversion), f32 (resample)",
            "data": "synthetic",
            "config": {"workload": wl["name"], "images_per_step_per_gpu": B,
                       "l2_policy": f"inputs rotate over {R} resident buffer sets; one batch is {by / 1e6:.0f} MB in+out > 126 MB L2",
                       "parallelism": f"images sharded by index over {world} GPU(s), no data-path collective",
                       "launch": ("one CUDA graph replay per step" if graphs is not None else "eager launches") +
                                 (f"; consecutive steps round-robin over {n_streams} CUDA streams" if ring is not None else "")},
            "clocks": clocks, "e2e": None, "gpu_launches": launches * world,
            "roofline": {"bound": "hbm", "kernel": "remap_f32_rows_kernel", "achieved": by / kms / 1e6, "peak": peak,
                         "unit": "GB/s", "frac": by / kms / 1e6 / peak, "traffic": ncu_traffic("remap_f32_rows_kernel", "c5"),
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": by, "kernel_ms": kms},
            "cpu_baseline": None, "per_rank_ms": [s[0] for s in stats]}
    print(json.dumps(line), flush=True)
    return 0


def run_gpu(args, rank, local_rank, world):
    wl = WORKLOADS[args.workload]
    if wl.get("ragged"):
        return run_gpu_ragged(args, rank, local_rank, world)
    if wl.get("pdf"):
        return run_gpu_pdf(args, rank, local_rank, world)
    cpu_info = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # before CUDA is initialised (the pool forks)
        cpu_info, _ = cpu_arm(args.workload, steps=3, warmup=1, budget_s=20.0, full_steps=False)

    import torch
    import torch.distributed as dist

    from attwarp_b200 import ops, sharding
    from attwarp_b200.batched import HostBatchPipeline, StreamRing

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        init_nccl(dev)

    B, side, C, grid = wl["B"], wl["side"], wl["C"], wl["grid"]
    L, Hh, T = wl["L"], wl["Hh"], wl["grid"] ** 2
    R = args.rotate if args.rotate > 0 else (8 if wl["has_attention"] else 4)
    gen = torch.Generator(device=dev).manual_seed(1235 + rank)
    sets = []
    for _ in range(R):
        s = {}
        if wl["has_attention"]:
            chunks = []
            for lo in range(0, B, 64):                 # bounded fp32 temporaries
                z = torch.randn(min(64, B - lo), L, Hh, T, device=dev, generator=gen) * 2
                chunks.append(torch.softmax(z, -1).to(torch.bfloat16))
            s["attn"] = torch.cat(chunks)
            del chunks, z
        else:
            tok = torch.rand(B, grid, grid, device=dev, generator=gen) ** 3
            s["tok"] = (tok / tok.sum(dim=(1, 2), keepdim=True)).contiguous()
        s["img"] = torch.randint(0, 256, (B, side, side, C), device=dev, dtype=torch.uint8, generator=gen)
        s["out"] = torch.empty(B, side, side, C, device=dev, dtype=torch.uint8)
        s["aux"] = (torch.empty(B, T, device=dev), torch.empty(B, side, device=dev),
                    torch.empty(B, side, device=dev))
        sets.append(s)

    n_steps_total = args.warmup + args.steps
    side_hw = (side, side)

    def enqueue(s, evs=None):
        """One step on buffer set `s`: every kernel of the hot path, on the current stream."""
        if wl["has_attention"]:
            ops.warp_from_attention_tokens(s["attn"], s["img"], (grid, grid), None, "hwc",
                                           transform="identity", out=s["out"], aux=s["aux"],
                                           stage_events=evs)
            return 3
        mx, my = s["aux"][1], s["aux"][2]
        ops.maps_from_tokens(s["tok"], side_hw, None, "identity", out=(mx, my))
        ops.remap_bilinear(s["img"], mx, my, "hwc", out=s["out"])
        return 2

    # one CUDA graph per resident buffer set: a step is ONE driver launch of its 2-3 kernels
    launches_per_step = 3 if wl["has_attention"] else 2
    graphs = None
    if not args.no_graph:
        graphs = [ops.GraphedCall(lambda s=s: enqueue(s), device=dev) for s in sets]

    # Consecutive steps are independent batches: they go round-robin over a few streams so that the
    # HBM-bound stage 1 of one step shares the GPU with the issue-bound stage 5 of the previous one.
    n_streams = args.streams if args.streams > 0 else (4 if wl["has_attention"] else 3)
    n_streams = max(1, min(n_streams, R))        # concurrent steps need distinct buffer sets
    ring = StreamRing(n_streams, dev) if n_streams > 1 else None

    def run_step(i):
        if graphs is not None:
            graphs[i % R].replay()
        else:
            enqueue(sets[i % R])

    def step(i):
        if ring is not None:
            ring.submit(lambda: run_step(i))
        else:
            run_step(i)
        return launches_per_step

    if ring is not None:
        ring.fork()
    for i in range(args.warmup):
        step(i)
    if ring is not None:
        ring.join()
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank, _gpu_uuid(local_rank))
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    ev0.record()
    if ring is not None:
        ring.fork()                               # every ring stream starts after ev0 ...
    for i in range(args.steps):
        launches += step(args.warmup + i)
    if ring is not None:
        ring.join()                               # ... and ev1 waits for all of them
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = ev0.elapsed_time(ev1)

    chk = sharding.checksum64(sets[(n_steps_total - 1) % R]["out"][:8])
    stats = sharding.gather_stats(elapsed_ms, B * args.steps, chk, dev)
    value = sharding.aggregate_throughput(stats)
    worst_ms = max(s[0] for s in stats)

    # ---- per-kernel durations, outside the timed region ---------------------------------------------
    # (a) stage breakdown of the fused call: CUDA events the library records between its kernels;
    # (b) the resample kernel alone, launched back to back over the rotating sets (the launch of
    #     kernel n+1 overlaps kernel n, so events around the whole run / N is the kernel duration).
    peak, peak_src = measured_peak()
    kernels = {}
    kreps = max(10, min(args.steps, 30))
    if wl["has_attention"]:
        stage_ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(kreps)]
        for evs in stage_ev:
            for e in evs:
                e.record()                              # materialise the cudaEvent_t handles
        for i in range(kreps):
            enqueue(sets[i % R], stage_ev[i])
        torch.cuda.synchronize()
        names = ["aggregate_rows_tma_kernel", "maps_from_tokens_kernel", "remap_u8_stream_kernel"]
        byts = [B * (L * Hh * T * 2 + T * 4), B * (T * 4 + 2 * side * 4), B * side * side * C * 2]
        for k, (nm, by) in enumerate(zip(names, byts)):
            ms = sum(evs[k].elapsed_time(evs[k + 1]) for evs in stage_ev) / kreps
            kernels[nm] = {"ms": ms, "algorithmic_bytes": by, "achieved_gbs": by / ms / 1e6,
                           "frac": by / ms / 1e6 / peak, "how": "library stage events inside the fused call"}
    def back_to_back(fn):
        """Average duration of one launch of a kernel launched kreps times in a row over the rotating
        sets (the launch of kernel n+1 overlaps kernel n, so there is no launch gap in the figure)."""
        for i in range(3):
            fn(sets[i % R])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(kreps):
            fn(sets[i % R])
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / kreps

    how = f"{kreps} back-to-back launches over the rotating sets"
    ms = back_to_back(lambda s: ops.remap_bilinear(s["img"], s["aux"][1], s["aux"][2], "hwc", out=s["out"]))
    by = B * side * side * C * 2
    kernels["remap_u8_stream_kernel"] = {"ms": ms, "algorithmic_bytes": by, "achieved_gbs": by / ms / 1e6,
                                        "frac": by / ms / 1e6 / peak, "how": how}
    if wl["has_attention"]:
        # stage 1 alone (the stage events above include the gap between the event and the kernel start)
        k1 = kernels["aggregate_rows_tma_kernel"]
        ms = back_to_back(lambda s: ops.aggregate_attention(s["attn"], out=s["aux"][0]))
        by = k1["algorithmic_bytes"]
        kernels["aggregate_rows_tma_kernel"] = {"ms": ms, "algorithmic_bytes": by, "achieved_gbs": by / ms / 1e6,
                                                "frac": by / ms / 1e6 / peak, "how": how,
                                                "ms_between_stage_events": k1["ms"]}
    dom = max(kernels, key=lambda k: kernels[k]["ms"])
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_gbs"], "peak": peak,
                "unit": "GB/s", "frac": kernels[dom]["frac"], "traffic": ncu_traffic(dom, args.workload),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": kernels[dom]["algorithmic_bytes"],
                "kernel_ms": kernels[dom]["ms"]}

    # ---- e2e: pinned host buffers -> H2D -> kernels -> D2H, per step -----------------------------
    e2e = None
    if wl["has_attention"] and not args.no_e2e:
        h_attn = torch.empty(B, L, Hh, T, dtype=torch.bfloat16).pin_memory()
        h_img = torch.empty(B, side, side, C, dtype=torch.uint8).pin_memory()
        h_out = torch.empty(B, side, side, C, dtype=torch.uint8).pin_memory()
        h_attn.copy_(sets[0]["attn"])
        h_img.copy_(sets[0]["img"])
        pipe = HostBatchPipeline(args.e2e_chunk, L, Hh, (grid, grid), (side, side, C), device=dev)
        for _ in range(2):
            pipe.run(h_attn, h_img, h_out)
        pipe.sync()
        ke = max(3, min(args.steps, 10))
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(ke):
            pipe.run(h_attn, h_img, h_out)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ems = e0.elapsed_time(e1)
        est = sharding.gather_stats(ems, B * ke, int(h_out[:4].to(torch.int64).sum()), dev)
        e2e = {"value": sharding.aggregate_throughput(est), "unit": UNIT,
               "h2d_bytes_per_step": h_attn.numel() * 2 + h_img.numel(),
               "d2h_bytes_per_step": h_out.numel(), "steps": ke, "chunk": args.e2e_chunk,
               "ms_per_step": max(s[0] for s in est) / ke,
               "api": "attwarp_b200.batched.HostBatchPipeline.run (pinned host in/out)"}

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": worst_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16->fp32 (stage 1), f64 (stages 2-4), u8 fixed-point (stage 5)",
            "data": "synthetic",
            "config": {"workload": wl["name"], "images_per_step_per_gpu": B,
                       "l2_policy": f"inputs rotate over {R} resident buffer sets; one batch "
                                    f"(attention+images) is {sum(t.numel() * t.element_size() for t in sets[0].values() if hasattr(t, 'numel')) / 1e6:.0f} MB > 126 MB L2",
                       "parallelism": f"images sharded by index over {world} GPU(s), no data-path collective",
                       "transform": "identity",
                       "launch": ("one CUDA graph replay per step" if graphs is not None else "eager launches") +
                                 (f"; consecutive steps round-robin over {n_streams} CUDA streams (independent batches: "
                                  "the early stages of one step overlap stage 5 of the previous one)" if ring is not None else "")},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches * world,
            "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu_info,
            "per_rank_ms": [s[0] for s in stats]}
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="default 200 (c4: 10)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--rotate", type=int, default=0, help="resident input buffer sets to rotate over (default 8 with attention, else 4)")
    ap.add_argument("--streams", type=int, default=0, help="CUDA streams consecutive steps alternate over (default 4 with attention, else 3)")
    ap.add_argument("--e2e-chunk", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the kernels eagerly instead of by graph replay")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 10 if args.workload == "c4" else 200
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank, local_rank, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    # stdout carries the ONE JSON line and nothing else: NCCL's own logging (its version banner, NCCL_DEBUG
    # output) goes to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if world == 1 and args.gpus > 1:
        print(f"bench.py: --gpus {args.gpus} needs torchrun (one rank per GPU); running 1 rank",
              file=sys.stderr)
    return run_gpu(args, rank, local_rank, world)


if __name__ == "__main__":
    sys.exit(main())
