"""GPU parity of the fused entry points at random batch / grid / image sizes: attention -> warped images in one call
(stages 1-5), the pinned-host pipeline around it, and PDFs -> warped images (BASELINE configs[4] chain), each against
the oracle chain or the unfused mirrors.  ATTWARP_FUZZ_CASES raises the number of cases (default 48)."""

import os

import numpy as np
import pytest
import torch

from gpu_util import need_gpu
from oracle import aggregate as OA
from oracle import numpy_path as ON
from oracle import torch_path as OT

pytestmark = pytest.mark.gpu

N_CASES = int(os.environ.get("ATTWARP_FUZZ_CASES", "48"))


@pytest.mark.parametrize("case", range(max(6, N_CASES // 4)))
def test_fused_attention_to_warp_random(case):
    need_gpu()
    from attwarp_b200 import batched, ops
    rng = np.random.default_rng(6000 + case)
    gen = torch.Generator().manual_seed(6000 + case)
    B, L, Hh = int(rng.integers(1, 10)), int(rng.integers(1, 5)), int(rng.integers(1, 9))
    gh, gw = int(rng.integers(2, 30)), int(rng.integers(2, 30))
    H, W = int(rng.integers(gh, 400)), int(rng.integers(gw, 700))
    Ho, Wo = int(rng.integers(2, 400)), int(rng.integers(2, 700))
    dt = [torch.bfloat16, torch.float16, torch.float32][case % 3]
    attn = torch.softmax(torch.randn(B, L, Hh, gh * gw, generator=gen) * 1.5, -1).to(dt)
    imgs = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)
    out, tok, mx, my = ops.warp_from_attention_tokens(attn.cuda(), torch.from_numpy(imgs).cuda(), (gh, gw), (Ho, Wo),
                                                      transform="identity", return_aux=True)
    torch.cuda.synchronize()
    tok_ref = OA.aggregate_attention(attn.float().numpy())
    assert np.max(np.abs(tok.cpu().numpy().reshape(B, -1) - tok_ref) / (tok_ref + 1e-12)) <= 1e-5
    tok_h = tok.cpu().numpy().reshape(B, gh, gw)
    for b in range(B):
        full = ON.upsample_tokens_nearest(tok_h[b], H, W)
        ref = ON.warp_image_by_attention(imgs[b], full, Wo, Ho, "identity")
        d = np.abs(out[b].cpu().numpy().astype(np.int32) - ref.astype(np.int32))
        assert d.max() <= 1 and (d != 0).mean() <= 5e-3, (case, b, (gh, gw), (H, W), (Ho, Wo), int(d.max()))
    # the same batch through the pinned-host pipeline, in chunks that do not divide it
    chunk = int(rng.integers(1, B + 2))
    pipe = batched.HostBatchPipeline(chunk, L, Hh, (gh, gw), (H, W, 3), (Ho, Wo), attn_dtype=dt)
    h_attn, h_img = attn.contiguous().pin_memory(), torch.from_numpy(imgs).pin_memory()
    h_out = torch.empty(B, Ho, Wo, 3, dtype=torch.uint8).pin_memory()
    pipe.run(h_attn, h_img, h_out)
    pipe.sync()
    assert torch.equal(h_out, out.cpu()), (case, "host pipeline differs from the device-resident call", chunk)


@pytest.mark.parametrize("case", range(max(6, N_CASES // 4)))
def test_pdfs_to_warp_random(case):
    need_gpu()
    from attwarp_b200 import checkpoint_utils as cu, model as mm, ops
    rng = np.random.default_rng(6500 + case)
    B, Nx, Ny = int(rng.integers(1, 9)), int(rng.integers(2, 40)), int(rng.integers(2, 40))
    C = int(rng.choice([1, 3, 4]))
    H, W = int(rng.integers(Ny, 300)), int(rng.integers(Nx, 500))
    Ho, Wo = int(rng.integers(2, 300)), int(rng.integers(2, 500))
    alpha = float(rng.choice([0.0, 0.1]))
    f32 = bool(case % 2)
    px = rng.random((B, Nx)).astype(np.float32) ** 3
    py = rng.random((B, Ny)).astype(np.float32) ** 3
    px /= px.sum(1, keepdims=True)
    py /= py.sum(1, keepdims=True)
    if f32:
        img = rng.random((B, C, H, W)).astype(np.float32)
    else:
        img = rng.integers(0, 256, (B, C, H, W), dtype=np.uint8)
    d_img, d_px, d_py = torch.from_numpy(img).cuda(), torch.from_numpy(px).cuda(), torch.from_numpy(py).cuda()
    out, Fx, Fy, mx, my = ops.warp_from_pdfs(d_img, d_px, d_py, alpha=alpha, out_size=(Ho, Wo), return_aux=True)
    # the unfused mirrors, step by step like trainer.py:212-218, 285-289: bit-identical
    Fx2 = cu.cdf_from_density(cu.upsample_pdf_right_inverse(mm.mix_with_uniform(d_px, alpha), W).clamp_min(0))
    Fy2 = cu.cdf_from_density(cu.upsample_pdf_right_inverse(mm.mix_with_uniform(d_py, alpha), H).clamp_min(0))
    out2 = cu.warp_from_cdf_torch(d_img, Fx2, Fy2, (Ho, Wo))
    torch.cuda.synchronize()
    assert torch.equal(Fx, Fx2) and torch.equal(Fy, Fy2) and torch.equal(out, out2), (case, "fused != unfused")
    # and against the oracle chain
    rFx = OT.cdf_from_density(np.maximum(OT.upsample_pdf_right_inverse(OT.mix_with_uniform(px, alpha), W), 0))
    assert np.abs(Fx.cpu().numpy() - rFx).max() <= 2e-5, (case, "Fx")
    rx, ry = OT.maps_from_cdf(Fx.cpu().numpy(), Fy.cpu().numpy(), (Ho, Wo))
    assert np.abs(mx.cpu().numpy() - rx).max() <= 1e-4 and np.abs(my.cpu().numpy() - ry).max() <= 1e-4
    ref = OT.warp_from_cdf(img[:1], Fx[:1].cpu().numpy(), Fy[:1].cpu().numpy(), (Ho, Wo))
    got = out[:1].cpu().numpy()
    if f32:
        dd = np.abs(got - ref)
        assert (dd > 1e-6).mean() <= 2e-3, (case, "f32 image", float(dd.max()))
    else:
        dd = np.abs(got.astype(np.int32) - ref.astype(np.int32))
        assert dd.max() <= 1 and (dd != 0).mean() <= 5e-3, (case, "u8 image", int(dd.max()))
