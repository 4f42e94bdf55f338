#!/usr/bin/env python
"""Generate golden vectors by EXECUTING THE UNMODIFIED REFERENCE (build container only).

    python tests/golden/make_golden.py

Imports the reference from /root/reference (via oracle/ref_loader.py, nothing is copied), runs
its hot-path functions on seeded synthetic inputs and stores inputs + outputs as compressed
``.npz`` files next to this script.  The reference has no tests or fixtures of its own
(SURVEY.md section 4), so these files are what pins both the oracle restatement
(tests/test_oracle_vs_golden.py, CPU) and the CUDA path (tests/test_gpu_*.py, GPU box --
where /root/reference does not exist).

Library versions at generation time are recorded in ``meta.json``.
"""

from __future__ import annotations

import contextlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import ref_loader as R  # noqa: E402

import torch  # noqa: E402
import cv2  # noqa: E402


def quiet():
    return contextlib.redirect_stderr(io.StringIO()), contextlib.redirect_stdout(io.StringIO())


class _Cv2Spy:
    """Proxy for the ``cv2`` module seen by new_method: records the maps handed to remap."""

    def __init__(self, real):
        self._real = real
        self.maps = None

    def __getattr__(self, k):
        return getattr(self._real, k)

    def remap(self, img, mx, my, **kw):
        self.maps = (np.array(mx[0, :], copy=True), np.array(my[:, 0], copy=True))
        return self._real.remap(img, mx, my, **kw)


def softmax_tokens(rng, scale, n=576):
    z = rng.standard_normal(n) * scale
    e = np.exp(z - z.max())
    return (e / e.sum()).astype(np.float32)


def smooth_image(h, w, c):
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    chans = [127 + 120 * np.sin(x / 9 + y / 13), 255 * x / max(w - 1, 1), 255 * y / max(h - 1, 1),
             127 + 120 * np.cos(x / 5 - y / 7)]
    img = np.stack(chans[:c], axis=-1)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def gen_numpy_path(rng):
    nm = R.new_method()
    spy = _Cv2Spy(cv2)
    nm.cv2 = spy
    cases = {}

    shared = {}

    def run(name, image, att, new_w, new_h, transform, exp_scale=1.0, exp_divisor=1.0,
            apply_inverse=False, image_key=None):
        nm.set_transform_function(transform, exp_scale, exp_divisor, apply_inverse)
        e, o = quiet()
        with e, o:
            out = nm.warp_image_by_attention(image, att, new_w, new_h)
        mx, my = spy.maps
        if image_key is not None:          # big images are stored once under shared/<key>
            shared[image_key] = image
            image = np.asarray(image_key)
        cases[name] = dict(image=image, att=att, new_w=new_w, new_h=new_h, transform=transform,
                           exp_scale=exp_scale, exp_divisor=exp_divisor,
                           apply_inverse=apply_inverse, out=out, map_x=mx, map_y=my)

    # C1: 336x336 RGB + 24x24 token map (BASELINE.json configs[0]), nearest-upsampled
    tok = softmax_tokens(rng, 2.0).reshape(24, 24)
    att336 = np.kron(tok, np.ones((14, 14), dtype=np.float32))
    noise336 = rng.integers(0, 256, (336, 336, 3), dtype=np.uint8)
    run("c1_noise_identity_336", noise336, att336, 336, 336, "identity", image_key="noise336")
    run("c1_noise_identity_500", noise336, att336, 500, 500, "identity", image_key="noise336")
    run("c1_smooth_sqrt_336", smooth_image(336, 336, 3), att336, 336, 336, "sqrt")
    # driver-style uint8 'mota mask' attention
    att_u8 = np.clip(np.rint(att336 / att336.max() * 255), 0, 255).astype(np.uint8)
    run("c1_smooth_u8att_identity_500", smooth_image(336, 336, 3), att_u8, 500, 500, "identity")

    # odd sizes, channel counts, transforms
    img = rng.integers(0, 256, (97, 53), dtype=np.uint8)
    run("odd_c1_sqrt", img, (rng.random((97, 53)) ** 3).astype(np.float32), 200, 64, "sqrt")
    img = rng.integers(0, 256, (61, 120, 4), dtype=np.uint8)
    run("odd_c4_square_inv", img, (rng.random((61, 120)) ** 3).astype(np.float32), 75, 130,
        "square", apply_inverse=True)
    img = rng.integers(0, 256, (64, 80, 3), dtype=np.uint8)
    attf = rng.random((64, 80)).astype(np.float32)
    run("exp_scaled", img, attf, 64, 70, "exp", 2.0, 3.0)
    run("exp_scaled_inv", img, attf, 90, 70, "exp", 2.0, 3.0, True)
    run("sqrt_inv", img, attf, 80, 64, "sqrt", apply_inverse=True)
    run("log_u8", img, rng.integers(0, 256, (64, 80), dtype=np.uint8), 64, 70, "log")
    run("f64_att", img, rng.random((64, 80)), 100, 100, "sqrt")
    run("unknown_transform", img, attf, 80, 64, "no-such-transform")

    # edge cases
    run("edge_all_zero_same_size", img, np.zeros((64, 80), np.uint8), 80, 64, "identity")
    run("edge_uniform_same_size", img, np.full((64, 80), 7, np.uint8), 80, 64, "identity")
    run("edge_uniform_upscale", img, np.full((64, 80), 1.0, np.float32), 160, 128, "sqrt")
    hot = np.zeros((64, 80), np.float32)
    hot[20, 30] = 1.0
    run("edge_single_hot", img, hot, 80, 64, "identity")
    run("edge_log_fallback_constant", img, (rng.random((64, 80)) * 0.5).astype(np.float32),
        80, 64, "log")
    run("edge_negative_att", img, (rng.standard_normal((64, 80))).astype(np.float32), 80, 64,
        "identity")

    flat = {f"shared/{k}": v for k, v in shared.items()}
    for name, d in cases.items():
        for k, v in d.items():
            flat[f"{name}/{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "numpy_path.npz"), **flat)
    return sorted(cases)


def gen_torch_path(rng):
    cu = R.checkpoint_utils()
    mm = R.marginalnet_model()
    out = {}
    t = torch.from_numpy

    logits = (rng.standard_normal((16, 24)) * 3).astype(np.float32)
    logits[3, 5] = np.nan
    logits[4, 2] = np.inf
    logits[5, 1] = -np.inf
    out["softmax/logits"] = logits
    p = mm.safe_softmax(t(logits)).numpy()
    out["softmax/p"] = p
    out["mix/alpha"] = np.float32(0.1)
    out["mix/p"] = mm.mix_with_uniform(t(p), 0.1).numpy()

    for L in (336, 512, 100):
        up = cu.upsample_pdf_right_inverse(t(p), L)
        out[f"upsample/{L}"] = up.numpy()
        out[f"upsample64/{L}"] = cu.upsample_pdf_right_inverse(t(p).double(), L).numpy()
        out[f"cdf_from_upsample/{L}"] = cu.cdf_from_density(up.clamp_min(0)).numpy()
    out["upsample/1d"] = cu.upsample_pdf_right_inverse(t(p[0]), 48).numpy()
    out["upsample/3d"] = cu.upsample_pdf_right_inverse(t(p.reshape(4, 4, 24)), 48).numpy()

    dens = rng.random((8, 300)).astype(np.float32)
    dens[0, 3] = np.nan
    dens[1, 4] = np.inf
    dens[2, :] = 0
    dens[3, 7] = -5
    out["cdf/p"] = dens
    out["cdf/F"] = cu.cdf_from_density(t(dens)).numpy()

    A = rng.standard_normal((4, 1, 97, 53)).astype(np.float32)
    mx, my = cu.gt_marginals(t(A))
    out["gt/A"], out["gt/mx"], out["gt/my"] = A, mx.numpy(), my.numpy()

    Af = rng.random((2, 1, 100, 130)).astype(np.float32)
    out["pool/A"] = Af
    out["pool/out"] = torch.nn.functional.adaptive_avg_pool2d(t(Af), (24, 24)).numpy()

    Fc = cu.cdf_from_density(t(rng.random((5, 24)).astype(np.float32)))
    out["resample/F"] = Fc.numpy()
    out["resample/out"] = cu.resample_cdf(Fc, 200).numpy()
    out["strict/out"] = cu._make_strictly_increasing(Fc).numpy()

    # warp_from_cdf_torch, incl. live tie-break branch (sharp PDFs) and out_size
    def pdf(scale, B):
        return mm.safe_softmax(t((rng.standard_normal((B, 24)) * scale).astype(np.float32)))

    specs = [("u8_same", 3, 3, 64, 80, None, np.uint8, 1.0),
             ("f32_out", 3, 3, 64, 80, (50, 70), np.float32, 4.0),
             ("u8_odd", 2, 3, 97, 53, (64, 200), np.uint8, 4.0),
             ("f32_c4", 2, 4, 120, 120, None, np.float32, 1.0),
             ("u8_c4_sharp", 2, 4, 100, 60, (100, 61), np.uint8, 6.0)]
    for name, B, C, H, W, osz, dt, scale in specs:
        if dt == np.uint8:
            img = rng.integers(0, 256, (B, C, H, W)).astype(np.uint8)
        else:
            img = rng.random((B, C, H, W)).astype(np.float32)
        Fx = cu.cdf_from_density(cu.upsample_pdf_right_inverse(pdf(scale, B), W).clamp_min(0))
        Fy = cu.cdf_from_density(cu.upsample_pdf_right_inverse(pdf(scale, B), H).clamp_min(0))
        res = cu.warp_from_cdf_torch(t(img), Fx, Fy, osz).numpy()
        out[f"warp/{name}/img"] = img
        out[f"warp/{name}/Fx"] = Fx.numpy()
        out[f"warp/{name}/Fy"] = Fy.numpy()
        out[f"warp/{name}/out_size"] = np.asarray([-1, -1] if osz is None else osz)
        out[f"warp/{name}/out"] = res
    np.savez_compressed(os.path.join(HERE, "torch_path.npz"), **out)
    return sorted(out)


def gen_aggregate(rng):
    lh = R.llava_hooks()
    out = {}
    B, L, Hh, K, T = 3, 4, 8, 700, 576
    logits = rng.standard_normal((B, L, Hh, K)).astype(np.float32)
    a = np.exp(logits)
    a /= a.sum(-1, keepdims=True)
    a_bf16 = torch.from_numpy(a).to(torch.bfloat16)
    starts = [1, 17, 64]
    out["attn_f32"] = a
    out["attn_bf16_bits"] = a_bf16.view(torch.int16).numpy()
    out["starts"] = np.asarray(starts)
    out["T"] = np.asarray(T)

    for tag, src in (("f32", torch.from_numpy(a)), ("bf16", a_bf16.float())):
        bl = lh.BatchMaskHookLogger(None, "cpu")
        bl.set_batch_image_token_ranges(starts, [s + T for s in starts])
        for l in range(L):
            q = torch.zeros(B, Hh, 2, K)
            q[:, :, -1, :] = src[:, l]
            bl._process_attention(q)
        out[f"batch_logger/{tag}"] = torch.stack(
            [m.reshape(-1) for m in bl.finalize_batch()]).numpy()
        ml = lh.MaskHookLogger(None, "cpu")
        ml.set_image_token_range(starts[1], starts[1] + T)
        for l in range(L):
            q = torch.zeros(1, Hh, 2, K)
            q[:, :, -1, :] = src[1:2, l]
            ml._process_attention(q)
        out[f"single_logger/{tag}"] = ml.finalize().numpy()
    # default range (no set_image_token_range): st=1, ed=min(577, kv)  llava.py:99-102
    ml = lh.MaskHookLogger(None, "cpu")
    q = torch.zeros(1, Hh, 3, K)
    q[:, :, -1, :] = torch.from_numpy(a[0:1, 0])
    ml._process_attention(q)
    out["single_logger/default_range"] = ml.finalize().numpy()
    # nothing captured -> uniform
    out["single_logger/empty"] = lh.MaskHookLogger(None, "cpu").finalize().numpy()
    # revise_mask on an aggregated map
    tm = torch.from_numpy(out["batch_logger/f32"][0].reshape(24, 24).copy())
    out["revise_mask/in"] = tm.numpy()
    out["revise_mask/out"] = lh.revise_mask(tm.float(), 3, 10).detach().numpy()
    np.savez_compressed(os.path.join(HERE, "aggregate.npz"), **out)
    return sorted(out)


def main():
    assert R.available(), "reference tree not found"
    rng = np.random.default_rng(20261017)
    torch.manual_seed(20261017)
    keys = {"numpy_path": gen_numpy_path(rng), "torch_path": gen_torch_path(rng),
            "aggregate": gen_aggregate(rng)}
    meta = {"numpy": np.__version__, "cv2": cv2.__version__, "torch": torch.__version__,
            "python": sys.version.split()[0], "cases": keys,
            "generator": "tests/golden/make_golden.py", "seed": 20261017}
    with open(os.path.join(HERE, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    for n in ("numpy_path", "torch_path", "aggregate"):
        print(n, os.path.getsize(os.path.join(HERE, n + ".npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
