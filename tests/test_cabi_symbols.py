"""CPU: the C-ABI library builds for sm_100a, loads without a GPU and exports every symbol that
include/attwarp.h declares (no compute calls are made here)."""

import os
import re

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "attwarp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(attwarp_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_loads_and_exports_all_symbols():
    from attwarp_b200 import _lib, build
    path = build.build()
    assert os.path.isfile(path)
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in attwarp.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in _lib.py"
    assert sorted(_lib.SIGNATURES) == declared
    assert lib.attwarp_abi_version() == 1


def test_argument_validation_without_gpu():
    """Invalid arguments are rejected before any CUDA call."""
    import ctypes as C
    from attwarp_b200 import _lib
    lib = _lib.load()
    rc = lib.attwarp_remap_bilinear(None, None, 0, 0, 1, 3, 4, 4, 4, 4, None, None, None)
    assert rc == _lib.ERR_INVALID_ARG
    assert b"NULL" in lib.attwarp_last_error()
    tp = _lib.make_transform("sqrt")
    tp.transform = 99
    buf = C.create_string_buffer(64)
    rc = lib.attwarp_maps_from_tokens(C.addressof(buf), 1, 2, 2, 4, 4, 4, 4, C.byref(tp),
                                      C.addressof(buf), C.addressof(buf), None)
    assert rc == _lib.ERR_INVALID_ARG and b"transform" in lib.attwarp_last_error()


def test_sass_is_sm100a():
    import shutil
    import subprocess
    from attwarp_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(cuobjdump):
        import pytest
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", _lib.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out
