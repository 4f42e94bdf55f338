#!/usr/bin/env python
"""Golden vectors for BASELINE configs[4] -- MarginalNet-predicted PDFs (hidden=256, batch 128) feeding
the CDF + resample path -- produced by EXECUTING THE UNMODIFIED REFERENCE (build container only).

    python tests/golden/make_golden_c5.py        ->  tests/golden/c5_marginalnet.npz

Chain executed (model/marginalnet_full_dataset/trainer.py:209-218, 285-289):
    px, py = MarginalNet(d_vis_in=1024, d_txt_in=4096, hidden=256)(fmap_v, 24, 24, txt_tok, txt_mask)
    p_s    = mix_with_uniform(p, alpha)                      alpha in {0, 0.1}
    p_img  = upsample_pdf_right_inverse(p_s, 512).clamp_min(0)
    F      = cdf_from_density(p_img)
    out    = warp_from_cdf_torch(img, Fx, Fy)                (B,3,512,512) float32 and uint8
Dv = 1024 / Dt = 4096 are the LLaVA-1.5 widths (the reference probes them at run time,
trainer.py:106-113); weights are random-initialised under a fixed seed, inputs are seeded noise.

Stored: px, py for all 128 samples; Fx, Fy for the first 32 (alpha = 0) and the first 8 (alpha = 0.1);
the warped images of the first 4 samples on every 8th row and column.  The input images are NOT stored:
`hash_image` below regenerates them bit-identically in the tests.
"""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

SEED = 1234 + 4          # seed = 1234 + config index
B, G, L = 128, 24, 512
N_CDF, N_CDF_MIX, N_WARP, STEP = 32, 8, 4, 8


def hash_image(n, c, h, w):
    """Deterministic uint8 noise-like image [n, c, h, w] from integer arithmetic only."""
    b = np.arange(n, dtype=np.int64)[:, None, None, None]
    ch = np.arange(c, dtype=np.int64)[None, :, None, None]
    y = np.arange(h, dtype=np.int64)[None, None, :, None]
    x = np.arange(w, dtype=np.int64)[None, None, None, :]
    v = (x * 73 + y * 151 + ch * 31 + b * 17) ^ ((x * y) >> 3) ^ ((x + 3 * y) * 2654435761 >> 7)
    return (v & 255).astype(np.uint8)


def main():
    import torch

    from oracle import ref_loader as R
    assert R.available(), "reference tree not found"
    cu, mm = R.checkpoint_utils(), R.marginalnet_model()
    torch.manual_seed(SEED)
    net = mm.MarginalNet(d_vis_in=1024, d_txt_in=4096, hidden=256).eval()
    gen = torch.Generator().manual_seed(SEED)
    fmap_v = torch.randn(B, 1024, G, G, generator=gen)
    txt_tok = torch.randn(B, 32, 4096, generator=gen)
    txt_mask = torch.ones(B, 32, 1)
    with torch.no_grad():
        px, py = net(fmap_v, G, G, txt_tok, txt_mask)
        # random-init heads give nearly flat PDFs; sharpen a copy so that the warp is not the identity
        # (the sharpened logits go back through the reference's own safe_softmax)
        px_sharp = mm.safe_softmax(torch.log(px) * 200.0)
        py_sharp = mm.safe_softmax(torch.log(py) * 200.0)
    out = {"px": px.numpy(), "py": py.numpy(), "px_sharp": px_sharp.numpy(), "py_sharp": py_sharp.numpy()}

    def cdfs(p, alpha):
        p_s = mm.mix_with_uniform(p, alpha)
        return cu.cdf_from_density(cu.upsample_pdf_right_inverse(p_s, L).clamp_min(0))

    Fx, Fy = cdfs(px_sharp, 0.0), cdfs(py_sharp, 0.0)
    out["Fx_a0"], out["Fy_a0"] = Fx[:N_CDF].numpy(), Fy[:N_CDF].numpy()
    out["Fx_a01"] = cdfs(px_sharp, 0.1)[:N_CDF_MIX].numpy()
    out["Fy_a01"] = cdfs(py_sharp, 0.1)[:N_CDF_MIX].numpy()
    out["Fx_flat"], out["Fy_flat"] = cdfs(px, 0.0)[:N_CDF_MIX].numpy(), cdfs(py, 0.0)[:N_CDF_MIX].numpy()
    img_u8 = hash_image(N_WARP, 3, L, L)
    img_f32 = img_u8.astype(np.float32) / np.float32(255.0)
    w_u8 = cu.warp_from_cdf_torch(torch.from_numpy(img_u8), Fx[:N_WARP], Fy[:N_WARP]).numpy()
    w_f32 = cu.warp_from_cdf_torch(torch.from_numpy(img_f32), Fx[:N_WARP], Fy[:N_WARP]).numpy()
    out["warp_u8_sub"] = np.ascontiguousarray(w_u8[:, :, ::STEP, ::STEP])
    out["warp_f32_sub"] = np.ascontiguousarray(w_f32[:, :, ::STEP, ::STEP])
    path = os.path.join(HERE, "c5_marginalnet.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB;",
          "identity-ness of the sharp warp (max |Fx - uniform|):",
          float(np.abs(Fx.numpy() - np.arange(1, L + 1) / L).max()))


if __name__ == "__main__":
    main()
