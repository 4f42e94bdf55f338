"""CPU: two structural guarantees of the product path.

1. Nothing under attwarp_b200/ imports, loads or executes anything under oracle/ (the oracle is test and
   bench infrastructure only); bench.py touches it only in the CPU arm, __graft_entry__ only in smoke().
2. A missing libattwarp_sm100.so fails loudly (ImportError with the build command) -- there is no CPU
   fallback to fall back to.
"""

import ast
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
PKG = os.path.join(ROOT, "attwarp_b200")


def _imports(path):
    tree = ast.parse(open(path).read(), path)
    names = []
    for node in ast.walk(tree):
        if isinstance(node, ast.Import):
            names += [a.name for a in node.names]
        elif isinstance(node, ast.ImportFrom):
            names.append(("." * node.level) + (node.module or ""))
    return names


def test_package_never_imports_the_oracle():
    for fn in sorted(os.listdir(PKG)):
        if fn.endswith(".py"):
            for name in _imports(os.path.join(PKG, fn)):
                assert not name.lstrip(".").startswith("oracle"), f"{fn} imports {name}"
            assert "oracle" not in open(os.path.join(PKG, fn)).read().replace("oracle port", ""), fn
    csrc = os.path.join(PKG, "csrc")
    for fn in sorted(os.listdir(csrc)):
        if fn.endswith((".cu", ".cuh", ".h")):
            assert "oracle/" not in open(os.path.join(csrc, fn)).read(), fn


def test_bench_and_entry_use_the_oracle_only_where_allowed():
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef):
            uses = [n for n in ast.walk(node) if isinstance(n, ast.ImportFrom) and (n.module or "").startswith("oracle")]
            if uses:
                assert node.name == "cpu_arm", f"bench.py:{node.name} imports the oracle"
    tree = ast.parse(open(os.path.join(ROOT, "__graft_entry__.py")).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef):
            uses = [n for n in ast.walk(node) if isinstance(n, ast.ImportFrom) and (n.module or "").startswith("oracle")]
            if uses and node.name == "build":
                # building the checker is not using it: build() may only run the baseline/_ref recipe
                assert all(n.module == "oracle" and [a.name for a in n.names] == ["make_ref"] for n in uses)
            elif uses:
                assert node.name == "smoke", f"__graft_entry__.py:{node.name} imports the oracle"


def test_missing_library_fails_loudly():
    code = (
        "import sys; sys.path.insert(0, r'%s')\n"
        "from attwarp_b200 import _lib\n"
        "_lib._LIB_PATH = r'/nonexistent/libattwarp_sm100.so'\n"
        "_lib._lib = None\n"
        "try:\n"
        "    _lib.load()\n"
        "except ImportError as e:\n"
        "    assert 'build' in str(e).lower(), str(e)\n"
        "    print('loud')\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == "loud"
