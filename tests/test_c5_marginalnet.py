"""BASELINE configs[4]: MarginalNet-predicted PDFs (hidden=256, batch 128) feeding the CDF + resample
path.  Golden vectors come from the unmodified reference (tests/golden/make_golden_c5.py).

CPU part: the oracle's restatement of the chain against the reference's outputs.
GPU part (-m gpu): the fused three-launch entry (ops.warp_from_pdfs / attwarp_warp_from_pdfs) at the
full configuration size against the golden CDFs, the oracle's maps and the golden warped images.
"""

import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import torch_path as OT
from golden.make_golden_c5 import B, L, N_CDF, N_CDF_MIX, N_WARP, STEP, hash_image


@pytest.fixture(scope="module")
def g5():
    return np.load(os.path.join(GOLDEN_DIR, "c5_marginalnet.npz"))


def _rel(a, b, floor=1e-8):
    return float(np.max(np.abs(a.astype(np.float64) - b) / (np.abs(b) + floor)))


def _oracle_cdf(p, alpha):
    return OT.cdf_from_density(np.maximum(OT.upsample_pdf_right_inverse(OT.mix_with_uniform(p, alpha), L), 0))


def test_oracle_chain_matches_reference(g5):
    assert g5["px"].shape == (B, 24) and abs(float(g5["px_sharp"][0].sum()) - 1.0) < 1e-5
    assert _rel(_oracle_cdf(g5["px_sharp"][:N_CDF], 0.0), g5["Fx_a0"]) <= 1e-5
    assert _rel(_oracle_cdf(g5["py_sharp"][:N_CDF], 0.0), g5["Fy_a0"]) <= 1e-5
    assert _rel(_oracle_cdf(g5["px_sharp"][:N_CDF_MIX], 0.1), g5["Fx_a01"]) <= 1e-5
    assert _rel(_oracle_cdf(g5["py"][:N_CDF_MIX], 0.0), g5["Fy_flat"]) <= 1e-5


def test_oracle_warp_matches_reference(g5):
    """The oracle's stage 4-5 restatement fed with the reference's CDFs: uint8 bit-equal, float32 1e-6."""
    img = hash_image(N_WARP, 3, L, L)
    Fx, Fy = g5["Fx_a0"][:N_WARP], g5["Fy_a0"][:N_WARP]
    out = OT.warp_from_cdf(img[:2], Fx[:2], Fy[:2])
    assert np.array_equal(out[:, :, ::STEP, ::STEP], g5["warp_u8_sub"][:2])
    outf = OT.warp_from_cdf(img[:1].astype(np.float32) / np.float32(255.0), Fx[:1], Fy[:1])
    assert np.abs(outf[:, :, ::STEP, ::STEP] - g5["warp_f32_sub"][:1]).max() <= 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["u8", "f32"])
def test_gpu_fused_pdfs_to_warp_full_size(g5, dtype):
    """B = 128, 3 x 512^2: CDFs within 1e-5 of the reference, maps within 1e-4 px of the oracle fed with
    the GPU's own CDFs, warped images against the golden subsample (uint8 +-1 LSB, float32 1/255 on at
    most 0.1 % of the pixels -- a coordinate on a 1/32-px rounding boundary may flip) and bit-equal to the
    unfused mirrors (checkpoint_utils.*) run one after the other."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from attwarp_b200 import checkpoint_utils as cu, model as mm, ops
    px, py = torch.from_numpy(g5["px_sharp"]).cuda(), torch.from_numpy(g5["py_sharp"]).cuda()
    img8 = np.concatenate([hash_image(N_WARP, 3, L, L)] * (B // N_WARP))       # 128 images, 4 distinct
    img = torch.from_numpy(img8).cuda()
    if dtype == "f32":
        img = img.float() / 255.0                     # float32 division, as in the generator
    out, Fx, Fy, mx, my = ops.warp_from_pdfs(img, px, py, alpha=0.0, return_aux=True)
    torch.cuda.synchronize()
    assert _rel(Fx[:N_CDF].cpu().numpy(), g5["Fx_a0"]) <= 1e-5
    assert _rel(Fy[:N_CDF].cpu().numpy(), g5["Fy_a0"]) <= 1e-5
    rx, ry = OT.maps_from_cdf(Fx[:8].cpu().numpy(), Fy[:8].cpu().numpy(), (L, L))
    assert np.abs(mx[:8].cpu().numpy() - rx).max() <= 1e-4 and np.abs(my[:8].cpu().numpy() - ry).max() <= 1e-4
    sub = out[:N_WARP, :, ::STEP, ::STEP].cpu().numpy()
    if dtype == "u8":
        d = np.abs(sub.astype(np.int32) - g5["warp_u8_sub"].astype(np.int32))
        assert d.max() <= 1 and (d != 0).mean() <= 1e-3
    else:
        d = np.abs(sub - g5["warp_f32_sub"])
        assert (d > 1e-6).mean() <= 1e-3 and d.max() <= 1.0 / 255.0 + 1e-6
    # the unfused mirrors, step by step like trainer.py:285-289
    Fx2 = cu.cdf_from_density(cu.upsample_pdf_right_inverse(px, L).clamp_min(0))
    Fy2 = cu.cdf_from_density(cu.upsample_pdf_right_inverse(py, L).clamp_min(0))
    assert torch.equal(Fx2, Fx) and torch.equal(Fy2, Fy)
    assert torch.equal(cu.warp_from_cdf_torch(img, Fx2, Fy2), out)
    # alpha-mix variant against the reference's CDFs
    _, Fxa, Fya, _, _ = ops.warp_from_pdfs(img[:N_CDF_MIX], px[:N_CDF_MIX], py[:N_CDF_MIX], alpha=0.1,
                                           return_aux=True)
    assert _rel(Fxa.cpu().numpy(), g5["Fx_a01"]) <= 1e-5 and _rel(Fya.cpu().numpy(), g5["Fy_a01"]) <= 1e-5
    assert torch.equal(cu.cdf_from_density(cu.upsample_pdf_right_inverse(
        mm.mix_with_uniform(px[:N_CDF_MIX], 0.1), L).clamp_min(0)), Fxa)
