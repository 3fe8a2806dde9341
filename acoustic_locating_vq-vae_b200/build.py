"""In-tree build of libb200vq.so: nvcc, sm_100a only, no torch involved.

`python -m` is not used (the package directory name carries a hyphen); call `build()` through
`__graft_entry__.build()` or run this file directly.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO_PATH = os.environ.get("B200VQ_SO") or os.path.join(CSRC, "libb200vq.so")   # B200VQ_SO: an experiment build (tools/)
SOURCES = ["b200vq.cu"]
HEADERS = ["common.cuh", "kernels_simt.cuh", "kernels_tc.cuh", "kernels_screen.cuh", "kernels_bwd.cuh", os.path.join("..", "..", "include", "b200vq.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libb200vq.so cannot be built (there is no CPU fallback)")


def needs_build() -> bool:
    if os.environ.get("B200VQ_SO"):       # an experiment build: use it as it is
        return False
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into csrc/libb200vq.so for sm_100a (cross-compiles without a GPU)."""
    if not force and not needs_build():
        return SO_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", SO_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return SO_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
