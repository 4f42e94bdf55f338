"""pytest configuration: markers, repo root on sys.path, golden-fixture loaders."""

import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


class Golden:
    """Groups the flat ``case/key`` entries of a golden .npz into per-case dicts."""

    def __init__(self, name):
        self.npz = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
        self.cases = {}
        for k in self.npz.files:
            case, _, key = k.rpartition("/")
            self.cases.setdefault(case, {})[key] = k

    def case(self, name):
        d = {k: self.npz[v] for k, v in self.cases[name].items()}
        if "image" in d and d["image"].dtype.kind in "US":       # reference to a shared image
            d["image"] = self.npz["shared/" + str(d["image"])]
        for k in ("new_w", "new_h"):
            if k in d:
                d[k] = int(d[k])
        for k in ("exp_scale", "exp_divisor"):
            if k in d:
                d[k] = float(d[k])
        if "apply_inverse" in d:
            d["apply_inverse"] = bool(d["apply_inverse"])
        if "transform" in d:
            d["transform"] = str(d["transform"])
        return d

    def names(self, prefix=""):
        return sorted(c for c in self.cases if c.startswith(prefix) and c != "shared")

    def __getitem__(self, key):
        return self.npz[key]


@pytest.fixture(scope="session")
def golden_numpy():
    return Golden("numpy_path")


@pytest.fixture(scope="session")
def golden_torch():
    return Golden("torch_path")


@pytest.fixture(scope="session")
def golden_aggregate():
    return Golden("aggregate")


def numpy_case_names():
    g = Golden("numpy_path")
    return [n for n in g.names() if n]
