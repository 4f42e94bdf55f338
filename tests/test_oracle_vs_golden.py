"""CPU: the oracle restatements reproduce the golden vectors recorded from the real reference
(tests/golden/make_golden.py), and the cv2.remap / np.interp restatements match the real
libraries.  Tolerances are those of BASELINE.md section 4."""

import numpy as np
import pytest

from oracle import aggregate as OA
from oracle import numpy_path as ON
from oracle import torch_path as OT
from conftest import numpy_case_names


def rel_err(a, b, floor=1e-8):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / (np.abs(b) + floor)))


# ---------------------------------------------------------------- numpy path, end to end
@pytest.mark.parametrize("name", numpy_case_names())
def test_numpy_path_bit_equal(golden_numpy, name):
    c = golden_numpy.case(name)
    out, mx, my = ON.warp_image_by_attention(
        c["image"], c["att"], c["new_w"], c["new_h"], ON.resolve_transform(c["transform"]),
        c["exp_scale"], c["exp_divisor"], c["apply_inverse"], return_maps=True)
    assert np.abs(mx - c["map_x"]).max() <= 1e-4          # grid coordinates, pixels
    assert np.abs(my - c["map_y"]).max() <= 1e-4
    assert out.shape == c["out"].shape and out.dtype == c["out"].dtype
    assert np.array_equal(out, c["out"])                  # 0 LSB


def test_edge_semantics(golden_numpy):
    c = golden_numpy.case("edge_all_zero_same_size")      # identity warp returns the input
    assert np.array_equal(c["out"], c["image"])
    c = golden_numpy.case("edge_uniform_same_size")
    assert np.array_equal(c["out"], c["image"])
    c = golden_numpy.case("edge_log_fallback_constant")   # documented constant-image quirk
    assert (c["out"] == c["image"][0, 0]).all()


# ---------------------------------------------------------------- third-party restatements
@pytest.mark.parametrize("shape", [(40, 50), (33, 71, 3), (64, 64, 4), (17, 19, 1)])
@pytest.mark.parametrize("dtype", [np.uint8, np.float32])
def test_remap_restatement_matches_cv2(shape, dtype):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(hash((shape, dtype.__name__)) % 2**32)
    H, W = shape[:2]
    img = (rng.integers(0, 256, shape).astype(np.uint8) if dtype == np.uint8
           else rng.random(shape).astype(np.float32))
    Ho, Wo = 57, 83
    # general (non-separable) float32 maps reaching past every border
    mx = (rng.random((Ho, Wo)) * (W + 6) - 3).astype(np.float32)
    my = (rng.random((Ho, Wo)) * (H + 6) - 3).astype(np.float32)
    ref = cv2.remap(img, mx, my, interpolation=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REPLICATE)
    mine = ON.remap(img, mx, my)
    if mine.ndim == 3 and mine.shape[2] == 1:
        mine = mine[..., 0]
    if dtype == np.uint8:
        assert np.array_equal(ref, mine)
    else:
        assert np.abs(ref - mine).max() <= 1e-6
    # separable maps (the only kind the warp produces), incl. exact .5/32 ties
    sx = np.sort(rng.random(Wo) * W).astype(np.float32)
    sy = np.sort(rng.random(Ho) * H).astype(np.float32)
    sx[:4] = [0.0, 1.0 / 64, 3.0 / 64, 0.5]
    ref = ON.remap_cv2(img, sx, sy)
    mine = ON.remap(img, sx, sy)
    if mine.ndim == 3 and mine.shape[2] == 1:
        mine = mine[..., 0]
    if dtype == np.uint8:
        assert np.array_equal(ref, mine)
    else:
        assert np.abs(ref - mine).max() <= 1e-6


def test_interp_restatement_matches_numpy():
    rng = np.random.default_rng(7)
    for n, m in [(5, 9), (337, 336), (337, 500), (1345, 1344), (54, 200), (98, 64)]:
        xp = np.concatenate(([0.0], np.cumsum(rng.random(n - 1) ** 4 + 1e-9)))
        xp = xp / xp[-1] * m
        xp[-1] = m
        x = np.arange(m, dtype=np.float32)
        ref = np.interp(x, xp, np.arange(n, dtype=np.float64))
        assert np.array_equal(ref, ON.interp_restated(x, xp))
        assert np.array_equal(ref, ON.interp_scalar_numpy_algorithm(x, xp))
    # repeated knots (ties) and the non-monotone fallback knots of SURVEY 7.3(ii)
    xp = np.array([0, 1, 1, 1, 2.5, 2.5, 4, 6.0])
    x = np.arange(6, dtype=np.float64)
    assert np.array_equal(np.interp(x, xp, np.arange(8.0)), ON.interp_restated(x, xp))
    xp = np.concatenate(([0.0], np.arange(1, 65) * 1e9 * 64))
    xp[-1] = 64
    x = np.arange(64, dtype=np.float64)
    assert np.array_equal(np.interp(x, xp, np.arange(65.0)), ON.interp_restated(x, xp))


# ---------------------------------------------------------------- torch path
def test_softmax_mix(golden_torch):
    g = golden_torch
    p = OT.safe_softmax(g["softmax/logits"])
    assert rel_err(p, g["softmax/p"]) <= 1e-5
    assert rel_err(OT.mix_with_uniform(g["softmax/p"], float(g["mix/alpha"])), g["mix/p"]) <= 1e-5


@pytest.mark.parametrize("L", [336, 512, 100])
def test_upsample_and_cdf(golden_torch, L):
    g = golden_torch
    up = OT.upsample_pdf_right_inverse(g["softmax/p"], L)
    # the reference solves in fp32 (LAPACK sgesv): its own noise vs its fp64 evaluation is up
    # to ~4e-7 absolute, so the tolerance carries that absolute floor (parity notes)
    assert np.abs(up - g[f"upsample/{L}"]).max() <= 1e-6
    assert np.abs(up - g[f"upsample64/{L}"]).max() <= 1e-6
    F = OT.cdf_from_density(np.maximum(g[f"upsample/{L}"], 0))
    assert rel_err(F, g[f"cdf_from_upsample/{L}"]) <= 1e-5


def test_upsample_shapes(golden_torch):
    g = golden_torch
    p = g["softmax/p"]
    assert np.abs(OT.upsample_pdf_right_inverse(p[0], 48) - g["upsample/1d"]).max() <= 1e-6
    assert np.abs(OT.upsample_pdf_right_inverse(p.reshape(4, 4, 24), 48)
                  - g["upsample/3d"]).max() <= 1e-6
    with pytest.raises(ValueError):
        OT.upsample_pdf_right_inverse(np.zeros((1, 1, 1, 24), np.float32), 48)


def test_cdf_marginals_pool_resample(golden_torch):
    g = golden_torch
    assert rel_err(OT.cdf_from_density(g["cdf/p"]), g["cdf/F"]) <= 1e-5
    mx, my = OT.gt_marginals(g["gt/A"])
    assert rel_err(mx, g["gt/mx"]) <= 1e-5 and rel_err(my, g["gt/my"]) <= 1e-5
    assert rel_err(OT.adaptive_avg_pool2d(g["pool/A"]), g["pool/out"]) <= 1e-5
    assert rel_err(OT.make_strictly_increasing(g["resample/F"]), g["strict/out"]) <= 1e-5
    assert rel_err(OT.resample_cdf(g["resample/F"], 200), g["resample/out"]) <= 1e-5


@pytest.mark.parametrize("name", ["u8_same", "f32_out", "u8_odd", "f32_c4", "u8_c4_sharp"])
def test_warp_from_cdf(golden_torch, name):
    g = golden_torch
    osz = tuple(int(v) for v in g[f"warp/{name}/out_size"])
    osz = None if osz[0] < 0 else osz
    out = OT.warp_from_cdf(g[f"warp/{name}/img"], g[f"warp/{name}/Fx"], g[f"warp/{name}/Fy"], osz)
    ref = g[f"warp/{name}/out"]
    assert out.shape == ref.shape and out.dtype == ref.dtype
    if ref.dtype == np.uint8:
        assert np.array_equal(out, ref)
    else:
        assert np.abs(out - ref).max() <= 1e-6


def test_warp_from_cdf_errors():
    img = np.zeros((1, 3, 8, 9), np.uint8)
    with pytest.raises(ValueError):
        OT.warp_from_cdf(img, np.zeros((1, 8), np.float32), np.zeros((1, 8), np.float32))
    with pytest.raises(AssertionError):
        OT.warp_from_cdf(img[0], np.zeros((1, 9), np.float32), np.zeros((1, 8), np.float32))


# ---------------------------------------------------------------- stage 1
def test_aggregate(golden_aggregate):
    g = golden_aggregate
    T = int(g["T"])
    starts = g["starts"]
    out = OA.aggregate_attention(g["attn_f32"], starts, T)
    assert rel_err(out, g["batch_logger/f32"]) <= 1e-5
    assert rel_err(out[1], g["single_logger/f32"]) <= 1e-5
    bits = g["attn_bf16_bits"].astype(np.uint16).astype(np.uint32) << 16
    a16 = bits.view(np.float32)
    assert rel_err(OA.aggregate_attention(a16, starts, T), g["batch_logger/bf16"]) <= 1e-5
    h = OA.HookAggregator()
    q = np.zeros((1, 8, 3, 700), np.float32)
    q[:, :, -1, :] = g["attn_f32"][0:1, 0]
    h.process(q)
    assert rel_err(h.finalize()[0], g["single_logger/default_range"]) <= 1e-5
    assert rel_err(OA.HookAggregator().finalize()[0], g["single_logger/empty"]) <= 1e-6
    assert np.abs(OA.revise_mask(g["revise_mask/in"]) - g["revise_mask/out"]).max() <= 1e-5
