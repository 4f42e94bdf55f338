"""CPU: the scalar device arithmetic in attwarp_b200/csrc/warp_math.h, compiled for the host by
g++ (tests/hostcheck, test-only), agrees with the oracle bit for bit.  This checks the formulas
the CUDA kernels use before any GPU time is spent; the kernels themselves are checked on the
GPU box by the ``-m gpu`` tests."""

import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import numpy_path as ON

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hc(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    out = tmp_path_factory.mktemp("hostcheck") / "libhostcheck.so"
    subprocess.check_call([gxx, "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(out),
                           os.path.join(HERE, "hostcheck", "hostcheck.cpp")])
    return ctypes.CDLL(str(out))


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_quantise(hc):
    rng = np.random.default_rng(0)
    m = np.concatenate([rng.random(5000).astype(np.float32) * 400 - 20,
                        (np.arange(-64, 640) / 64.0).astype(np.float32)])   # exact .5 ties
    s = np.empty(m.size, np.int32)
    hc.hc_quantise(_p(m), m.size, _p(s))
    i, a = ON.quantise_coord(m)
    assert np.array_equal(s.astype(np.int64) >> 5, i) and np.array_equal(s & 31, a)


@pytest.mark.parametrize("C", [1, 3, 4])
def test_remap(hc, C):
    rng = np.random.default_rng(C)
    H, W, Ho, Wo = 37, 53, 41, 67
    mx = np.sort(rng.random(Wo) * (W + 2) - 1).astype(np.float32)
    my = np.sort(rng.random(Ho) * (H + 2) - 1).astype(np.float32)
    img = rng.integers(0, 256, (H, W, C), dtype=np.uint8)
    out = np.empty((Ho, Wo, C), np.uint8)
    hc.hc_remap_u8(_p(img), _p(out), C, H, W, Ho, Wo, _p(mx), _p(my))
    assert np.array_equal(out, ON.remap_u8(img, mx, my))
    imgf = rng.random((H, W, C)).astype(np.float32)
    outf = np.empty((Ho, Wo, C), np.float32)
    hc.hc_remap_f32(_p(imgf), _p(outf), C, H, W, Ho, Wo, _p(mx), _p(my))
    assert np.array_equal(outf, ON.remap_f32(imgf, mx, my))


def test_interp(hc):
    rng = np.random.default_rng(3)
    for n, m in [(5, 9), (337, 336), (337, 500), (1345, 1344), (54, 200), (98, 64)]:
        xp = np.concatenate(([0.0], np.cumsum(rng.random(n - 1) ** 4 + 1e-9)))
        xp = xp / xp[-1] * m
        xp[-1] = m
        out = np.empty(m, np.float64)
        hc.hc_interp(_p(xp), n, m, _p(out))
        assert np.array_equal(out, np.interp(np.arange(m), xp, np.arange(n, dtype=np.float64)))
    xp = np.array([0, 1, 1, 1, 2.5, 2.5, 4, 6.0])
    out = np.empty(6, np.float64)
    hc.hc_interp(_p(xp), 8, 6, _p(out))
    assert np.array_equal(out, np.interp(np.arange(6), xp, np.arange(8.0)))
    xp = np.concatenate(([0.0], np.arange(1, 65) * 1e9 * 64))     # fallback quirk knots
    xp[-1] = 64
    out = np.empty(64, np.float64)
    hc.hc_interp(_p(xp), 65, 64, _p(out))
    assert np.array_equal(out, np.interp(np.arange(64), xp, np.arange(65.0)))


def test_transforms(hc):
    rng = np.random.default_rng(4)
    x = np.concatenate([rng.random(1000) * 255, [0.0, 1e-12, 255.0]])
    for t, name in enumerate(ON.TRANSFORM_NAMES):
        for inv in (0, 1):
            out = np.empty_like(x)
            hc.hc_transform(_p(x), x.size, t, ctypes.c_double(2.0), ctypes.c_double(3.0), inv, _p(out))
            xin = x / 255.0 if name == "exp" else x
            if name == "exp":
                hc.hc_transform(_p(np.ascontiguousarray(xin)), x.size, t, ctypes.c_double(2.0),
                                ctypes.c_double(3.0), inv, _p(out))
            ref = (ON.inverse_transform if inv else ON.forward_transform)(xin, name, 2.0, 3.0)
            np.testing.assert_allclose(out, ref, rtol=1e-15, atol=0)
