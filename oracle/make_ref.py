#!/usr/bin/env python
"""Recipe for ``baseline/_ref/``: the UNMODIFIED reference files of the hot path, so that they can travel to the
GPU box (``/root/reference`` does not exist there; ``baseline/_ref/`` is git-ignored but not gpurun-ignored).

    python oracle/make_ref.py            (build container only; also run by __graft_entry__.build())

TEST / BENCH INFRASTRUCTURE ONLY.  The reference is pure Python with no installable package
(``pip install --target baseline/_ref /root/reference`` fails: no setup.py / pyproject.toml), so the
files of SURVEY.md section 8(c) are placed byte for byte, in the reference's own relative layout, under
``baseline/_ref/`` and a MANIFEST.json records their SHA-256.  ``oracle/ref_loader.py`` imports them from there
when ``/root/reference`` is absent; ``bench.py --impl reference`` then times the real reference functions
(``cpu_baseline.kind = "reference"``) instead of the oracle port.  Nothing under ``baseline/_ref`` is part of the
repository history or of the product.
"""

from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
SRC = os.environ.get("ATTWARP_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")

FILES = [
    "Attention Guided Warping/new_method.py",
    "Attention Guided Warping/attention_extraction/__init__.py",
    "Attention Guided Warping/attention_extraction/functions.py",      # imported by the package __init__
    "Attention Guided Warping/attention_extraction/llava.py",
    "model/marginalnet_full_dataset/checkpoint_utils.py",
    "model/marginalnet_full_dataset/model.py",
]


def make(verbose: bool = True) -> bool:
    """Place the files; returns False (and leaves an existing copy alone) when the source tree is absent."""
    if not os.path.isfile(os.path.join(SRC, FILES[0])):
        if verbose:
            print(f"oracle/make_ref.py: {SRC} not present; keeping {DST} as it is")
        return os.path.isfile(os.path.join(DST, FILES[0]))
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "sha256": manifest}, f, indent=1)
    if verbose:
        print(f"oracle/make_ref.py: {len(FILES)} reference files -> {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
