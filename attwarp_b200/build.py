"""Build libattwarp_sm100.so in-tree with nvcc for sm_100a (no GPU needed to compile).

    python -m attwarp_b200.build [--force] [--verbose]

One translation unit per .cu file, compiled in parallel, linked into
``attwarp_b200/libattwarp_sm100.so`` (git-ignored; it travels to the GPU box with the
snapshot).  Objects are cached under ``attwarp_b200/csrc/build/`` and rebuilt when a source or
header is newer.
"""

from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.abspath(os.path.join(PKG_DIR, "..", "include"))
OBJ_DIR = os.path.join(CSRC, "build")
LIB_PATH = os.path.join(PKG_DIR, "libattwarp_sm100.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
    "--expt-relaxed-constexpr",
    "-I", INCLUDE,
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hs += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE) if f.endswith(".h")]
    return hs


def _newer(target: str, deps) -> bool:
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("command failed:\n" + " ".join(cmd) + "\n" + res.stdout)
    if verbose and res.stdout.strip():
        print(res.stdout)
    return res.stdout


def build(force: bool = False, verbose: bool = False, ptxas_info: bool = False) -> str:
    nvcc = nvcc_path()
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdrs = headers() + [os.path.abspath(__file__)]
    jobs = []
    objs = []
    for src in sources():
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + hdrs):
            extra = os.environ.get("ATTWARP_NVCC_EXTRA", "").split()   # tuning experiments (-D...)
            cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if ptxas_info else []) + ["-c", src, "-o", obj]
            jobs.append(cmd)
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            outs = list(ex.map(lambda c: _run(c, verbose), jobs))
        if ptxas_info:
            print("\n".join(outs))
    if force or jobs or _newer(LIB_PATH, objs):
        _run([nvcc, "-shared", "-o", LIB_PATH] + objs +
             ["-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"], verbose)
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv,
                 ptxas_info="--ptxas" in sys.argv)
    print(path)
