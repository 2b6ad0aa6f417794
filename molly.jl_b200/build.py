"""Build libmollyb200.so in-tree with nvcc for sm_100a (the only target)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmollyb200.so")
SOURCES = ["engine.cu"]
# every header under csrc/ plus the C ABI header: an edit to any of them makes the library stale
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [os.path.join("..", "..", "include", "mollyb200.h")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libmollyb200 cannot be built (there is no CPU fallback)")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in SOURCES + HEADERS:
        if os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    return False


def build(force: bool = False, verbose: bool = False, defines=(), out: str = None) -> str:
    """defines/out: build a tuning variant (e.g. defines=("MB_LIST_BATCH=4",), out="libmollyb200_b4.so")."""
    if out is None and not force and not needs_build():
        return LIB
    host_cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [
        _nvcc(), "-std=c++17", "-O3", "-lineinfo",
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-ccbin", host_cxx,
        "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function",
        "--expt-relaxed-constexpr",
        "-shared", "-o", os.path.join(HERE, out) if out else LIB,
    ]
    cmd += [f"-D{d}" for d in defines]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    cmd += ["-lcudart"]
    print("[mollyb200] " + " ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return os.path.join(HERE, out) if out else LIB


if __name__ == "__main__":
    defs = tuple(a[2:] for a in sys.argv[1:] if a.startswith("-D"))
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=outs[0] if outs else None)
