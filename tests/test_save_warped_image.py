"""``save_warped_image`` -- the wrapper the reference drivers call (AGW/new_method.py:405-506; main.py:520,
main_batched.py:280) -- against outputs of the UNMODIFIED reference (tests/golden/save_warped.npz, written by
tests/golden/make_golden_save.py): the decoded PNG the reference wrote for a PIL image + image-size uint8 mask
(the drivers' call), for a 24 x 24 attention map (the image is shrunk to the map's size before warping,
new_method.py:478), for a path / list / 3-D / PIL attention input and for a grey image.

CPU: the oracle's array-level restatement reproduces the reference's PNG bit for bit.
GPU (-m gpu): the drop-in writes the same PNG (uint8 images: +-1 LSB stated, 0 measured) and returns True.
"""

import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import numpy_path as ON

CASES = ["driver_336_to_500", "quirk_24x24", "path_list_sqrt", "att_3d_mean", "pil_att_exp_inv", "gray_input"]


@pytest.fixture(scope="module")
def gs():
    return np.load(os.path.join(GOLDEN_DIR, "save_warped.npz"))


def _args(gs, name):
    return dict(width=int(gs[name + "/width"]), height=int(gs[name + "/height"]), transform=str(gs[name + "/transform"]),
                exp_scale=float(gs[name + "/exp_scale"]), exp_divisor=float(gs[name + "/exp_divisor"]),
                apply_inverse=bool(gs[name + "/apply_inverse"]))


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_png(gs, name):
    assert bool(gs[name + "/ok"])
    got = ON.save_warped_image_arrays(gs[name + "/image_rgb"], gs[name + "/att"], **_args(gs, name))
    ref = gs[name + "/warped_bgr"]
    assert got.shape == ref.shape and np.array_equal(got, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_save_warped_image_matches_reference(gs, name, tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import cv2
    from PIL import Image
    from attwarp_b200 import new_method as NM
    image, att, kw = gs[name + "/image_rgb"], gs[name + "/att"], _args(gs, name)
    p_orig, p_ov, p_out = (str(tmp_path / n) for n in ("orig.png", "overlay.png", "warped.png"))
    p_vis = str(tmp_path / "vis.png") if name == "path_list_sqrt" else None
    if name == "path_list_sqrt":                       # image by path, attention as a one-element list
        src = str(tmp_path / "input.png")
        cv2.imwrite(src, cv2.cvtColor(image, cv2.COLOR_RGB2BGR))
        img_arg, att_arg = src, [att]
    elif name == "pil_att_exp_inv":
        img_arg, att_arg = Image.fromarray(image), Image.fromarray(att, mode="L")
    else:
        img_arg, att_arg = Image.fromarray(image), att
    ok = NM.save_warped_image(img_arg, att_arg, p_orig, p_ov, p_out, p_vis, kw["width"], kw["height"], kw["transform"],
                              kw["exp_scale"], kw["exp_divisor"], kw["apply_inverse"])
    assert ok is True
    ref = gs[name + "/warped_bgr"]
    got = cv2.imread(p_out, cv2.IMREAD_UNCHANGED)
    assert got.shape == ref.shape == (kw["height"], kw["width"], 3)
    diff = np.abs(got.astype(int) - ref.astype(int))
    assert diff.max() <= 1, f"{name}: {diff.max()} LSB"                 # uint8 +-1 LSB
    assert (diff != 0).mean() <= 1e-3
    # the copy of the input and the overlay are host-side cv2 work, identical calls -> identical files
    want = image if image.ndim == 3 else np.repeat(image[..., None], 3, -1)
    assert np.array_equal(cv2.imread(p_orig, cv2.IMREAD_UNCHANGED), want[..., ::-1])
    ov = cv2.imread(p_ov, cv2.IMREAD_UNCHANGED)
    if name + "/overlay_bgr" in gs.files:
        assert np.array_equal(ov, gs[name + "/overlay_bgr"])
    else:
        assert tuple(ov.shape) == tuple(gs[name + "/overlay_shape"])
    if p_vis:
        vis = cv2.imread(p_vis, cv2.IMREAD_UNCHANGED)
        assert vis is not None and vis.ndim == 3 and vis.shape[0] == max(att.shape[0], kw["height"])
    # the module-level transform state is what the reference leaves behind (new_method.py:483)
    assert NM.ATTENTION_TRANSFORM == kw["transform"]


@pytest.mark.gpu
def test_gpu_save_warped_image_empty_list_and_failure(tmp_path):
    """An empty attention list becomes a constant 128 map of (height, width) (new_method.py:445-447): the image is
    resized to it and the uniform map is the identity warp; an unreadable path returns False without raising."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import cv2
    from PIL import Image
    from attwarp_b200 import new_method as NM
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (60, 90, 3), dtype=np.uint8)
    out = str(tmp_path / "w.png")
    assert NM.save_warped_image(Image.fromarray(img), [], None, None, out, None, 48, 36) is True
    want = cv2.resize(np.ascontiguousarray(img[..., ::-1]), (48, 36), interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(cv2.imread(out, cv2.IMREAD_UNCHANGED), want)
    assert NM.save_warped_image(str(tmp_path / "missing.jpg"), np.ones((8, 8)), None, None, out) is False
