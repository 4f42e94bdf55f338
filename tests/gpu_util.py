"""Helpers shared by the ``-m gpu`` parity tests."""

import numpy as np
import pytest
import torch


def need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def rel_err(a, b, floor=1e-8):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / (np.abs(b) + floor)))


def hwc(img):
    return img if img.ndim == 3 else img[..., None]
