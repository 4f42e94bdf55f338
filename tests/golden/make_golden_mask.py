#!/usr/bin/env python
"""Golden vectors for the mask post-processing, produced by EXECUTING THE
UNMODIFIED REFERENCE ``blend_mask`` (llava.py:240-270) in the build container.

    python tests/golden/make_golden_mask.py      ->  tests/golden/mask_path.npz

For every case: the [24, 24] token map, the uint8 24 x 24 mask after ``revise_mask`` + ``ToPILImage``
(recomputed here with the reference's own functions), and the mode-'L' mask ``blend_mask`` returns at the
image size.
"""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

SEED = 1234 + 0
CASES = [("c1_336", 336, 336, 2.0), ("wide_500x333", 333, 500, 1.0), ("tall_97x53", 97, 53, 3.0),
         ("big_1344", 1344, 1344, 2.0), ("small_20x30", 20, 30, 1.0)]     # name, H, W, softmax scale


def main():
    import torch
    from PIL import Image

    from oracle import ref_loader as R
    assert R.available(), "reference tree not found"
    lh = R.llava_hooks()
    rng = np.random.default_rng(SEED)
    out = {}
    for name, H, W, scale in CASES:
        z = rng.standard_normal(576) * scale
        tok = (np.exp(z - z.max()) / np.exp(z - z.max()).sum()).astype(np.float32).reshape(24, 24)
        image = Image.fromarray(rng.integers(0, 256, (H, W, 3), dtype=np.uint8), mode="RGB")
        _, mask = lh.blend_mask(image, torch.from_numpy(tok), 10, 3, Image.LANCZOS, 0)
        rev = lh.revise_mask(torch.from_numpy(tok).float(), kernel_size=3, enhance_coe=10)
        u8 = np.array(lh.toImg(rev.detach().reshape(1, 24, 24)))
        out[f"{name}/tok"] = tok
        out[f"{name}/revised"] = rev.detach().numpy().reshape(24, 24)
        out[f"{name}/u8_24"] = u8
        out[f"{name}/mask"] = np.array(mask.convert("L"))
        assert out[f"{name}/mask"].shape == (H, W)
    path = os.path.join(HERE, "mask_path.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
