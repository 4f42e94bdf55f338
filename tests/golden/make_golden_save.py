#!/usr/bin/env python
"""Golden vectors for the wrapper the reference drivers call: ``save_warped_image``
("Attention Guided Warping/new_method.py":405-506; call sites main.py:520, main_batched.py:280).

    python tests/golden/make_golden_save.py            (build container only: imports /root/reference)

Runs the UNMODIFIED reference function on seeded inputs, reads back the PNG files it wrote and stores the
decoded arrays (inputs + outputs) in ``save_warped.npz``.  Cases:

  driver_336_to_500   PIL RGB image + the uint8 image-size mask ``blend_mask`` returns, 500 x 500,
                      "identity" -- exactly what main.py:520-533 passes
  quirk_24x24         PIL image + a 24 x 24 float32 attention map: the reference shrinks the IMAGE to
                      24 x 24 with cv2.resize(INTER_LINEAR) before warping (new_method.py:478)
  path_list_sqrt      image given as a file path, attention as a one-element list, "sqrt", vis strip on
  att_3d_mean         an [h, w, 3] attention map is averaged over its channel axis (new_method.py:449-450)
  pil_att_exp_inv     attention given as a PIL 'L' image, transform "exp" with scale/divisor + apply_inverse
  gray_input          a 2-D (mode 'L') PIL image: cv2.cvtColor(RGB2BGR) of this OpenCV build replicates the single
                      channel, so the wrapper warps a grey BGR image
The reference's copy of the input (``original_image_save_path``) is checked here to equal the input and is not
stored; the JET overlay is stored for the small cases only.
"""

from __future__ import annotations

import contextlib
import io
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import ref_loader as R  # noqa: E402

import cv2  # noqa: E402
from PIL import Image  # noqa: E402


def smooth_noise(rng, h, w):
    """Half smooth, half noise: compresses a little and still exercises every rounding path."""
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    base = np.stack([127 + 120 * np.sin(x / 9 + y / 13), 255 * x / max(w - 1, 1), 127 + 120 * np.cos(x / 5 - y / 7)], -1)
    img = np.clip(np.rint(base), 0, 255).astype(np.uint8)
    noise = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    img[:, w // 2:] = noise[:, w // 2:]
    return img


def main():
    assert R.available(), "reference tree not found"
    nm = R.new_method()
    rng = np.random.default_rng(20261018)
    out = {}

    def run(name, image, att, width, height, transform="identity", exp_scale=1.0, exp_divisor=1.0,
            apply_inverse=False, as_path=False, vis=False, att_wrap=None):
        with tempfile.TemporaryDirectory() as d:
            p_orig, p_ov, p_out = (os.path.join(d, n) for n in ("orig.png", "overlay.png", "warped.png"))
            p_vis = os.path.join(d, "vis.png") if vis else None
            if as_path:
                src = os.path.join(d, "input.png")
                cv2.imwrite(src, cv2.cvtColor(image, cv2.COLOR_RGB2BGR))          # lossless: imread gives it back
                img_arg = src
            else:
                img_arg = Image.fromarray(image)
            att_arg = att if att_wrap is None else att_wrap(att)
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf), contextlib.redirect_stderr(buf):
                ok = nm.save_warped_image(img_arg, att_arg, p_orig, p_ov, p_out, p_vis, width, height, transform,
                                          exp_scale, exp_divisor, apply_inverse)
            out[f"{name}/image_rgb"] = image
            out[f"{name}/att"] = np.asarray(att)
            out[f"{name}/width"], out[f"{name}/height"] = np.asarray(width), np.asarray(height)
            out[f"{name}/transform"] = np.asarray(transform)
            out[f"{name}/exp_scale"], out[f"{name}/exp_divisor"] = np.asarray(exp_scale), np.asarray(exp_divisor)
            out[f"{name}/apply_inverse"] = np.asarray(apply_inverse)
            out[f"{name}/ok"] = np.asarray(bool(ok))
            if ok:
                out[f"{name}/warped_bgr"] = cv2.imread(p_out, cv2.IMREAD_UNCHANGED)
                orig = cv2.imread(p_orig, cv2.IMREAD_UNCHANGED)
                want = image if image.ndim == 3 else np.repeat(image[..., None], 3, -1)
                assert np.array_equal(orig, want[..., ::-1]), name
                ov = cv2.imread(p_ov, cv2.IMREAD_UNCHANGED)
                if ov.size <= 60000:
                    out[f"{name}/overlay_bgr"] = ov
                else:
                    out[f"{name}/overlay_shape"] = np.asarray(ov.shape)
                if vis:
                    out[f"{name}/vis_shape"] = np.asarray(cv2.imread(p_vis, cv2.IMREAD_UNCHANGED).shape)
            else:
                assert not os.path.exists(p_out)

    # main.py-style: uint8 mask at image size (a smooth bump field like a LANCZOS-upsampled 24 x 24 mask)
    tok = rng.random((24, 24)) ** 3
    mask = np.asarray(Image.fromarray(np.clip(tok / tok.max() * 255, 0, 255).astype(np.uint8), mode="L")
                      .resize((336, 336), Image.LANCZOS))
    run("driver_336_to_500", smooth_noise(rng, 336, 336), mask, 500, 500)
    run("quirk_24x24", smooth_noise(rng, 200, 300), (rng.random((24, 24)) ** 2).astype(np.float32), 64, 48)
    att = (rng.random((120, 160)) ** 3).astype(np.float32)
    run("path_list_sqrt", smooth_noise(rng, 120, 160), att, 200, 150, "sqrt", as_path=True, vis=True,
        att_wrap=lambda a: [a])
    run("att_3d_mean", smooth_noise(rng, 90, 70), rng.integers(0, 256, (90, 70, 3), dtype=np.uint8), 70, 90)
    run("pil_att_exp_inv", smooth_noise(rng, 64, 80), rng.integers(0, 256, (64, 80), dtype=np.uint8), 100, 60,
        "exp", 2.0, 3.0, True, att_wrap=lambda a: Image.fromarray(a, mode="L"))
    gray = smooth_noise(rng, 40, 50)[..., 0]
    run("gray_input", gray, rng.random((40, 50)).astype(np.float32), 50, 40)

    path = os.path.join(HERE, "save_warped.npz")
    np.savez_compressed(path, **out)
    print("save_warped", os.path.getsize(path) // 1024, "KiB;", {k.split("/")[0]: bool(out[k]) for k in out if k.endswith("/ok")})


if __name__ == "__main__":
    main()
