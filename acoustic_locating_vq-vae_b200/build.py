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
    "-Xlinker", "-soname=libb200vq.so",      # the torch binding links against it: one instance per process
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


DEBUG_SO_PATH = os.path.join(CSRC, "libb200vq_debug.so")


def build(force: bool = False, verbose: bool = False, debug: bool = False) -> str:
    """Compile csrc/*.cu into csrc/libb200vq.so for sm_100a (cross-compiles without a GPU).
    debug=True builds csrc/libb200vq_debug.so with -DVQ_DEBUG (device-side bounds checks, common.cuh); run a test with it as
    `B200VQ_SO=<that path> B200VQ_PY_AUTOGRAD=1 python -m pytest tests/test_gpu_properties.py -m gpu`."""
    if debug:
        cmd = [_nvcc(), *NVCC_FLAGS, "-DVQ_DEBUG", "-o", DEBUG_SO_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
        return DEBUG_SO_PATH
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


# ---- the C++ autograd node of the nn.Module (torch_binding.cpp): built in-tree with torch's extension machinery ----------
TORCH_EXT_NAME = "b200vq_torch"
TORCH_EXT_DIR = os.path.join(CSRC, "_torch")
TORCH_EXT_SO = os.path.join(TORCH_EXT_DIR, TORCH_EXT_NAME + ".so")


def torch_binding_needs_build() -> bool:
    if not os.path.exists(TORCH_EXT_SO):
        return True
    t = os.path.getmtime(TORCH_EXT_SO)
    deps = [os.path.join(CSRC, "torch_binding.cpp"), os.path.join(HERE, "..", "include", "b200vq.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_torch_binding(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/torch_binding.cpp into csrc/_torch/b200vq_torch.so (links libb200vq.so; needs no GPU)."""
    build(force=False)
    if not force and not torch_binding_needs_build():
        return TORCH_EXT_SO
    import ctypes
    from torch.utils import cpp_extension
    os.makedirs(TORCH_EXT_DIR, exist_ok=True)
    # load() imports what it built: libb200vq.so must already be in the process (found by its soname, as at run time)
    ctypes.CDLL(SO_PATH, mode=ctypes.RTLD_GLOBAL)
    cpp_extension.load(name=TORCH_EXT_NAME, sources=[os.path.join(CSRC, "torch_binding.cpp")], build_directory=TORCH_EXT_DIR,
                       extra_cflags=["-O2", "-std=c++17"], extra_ldflags=[f"-L{CSRC}", "-l:libb200vq.so", f"-Wl,-rpath,{CSRC}"],
                       with_cuda=True, verbose=verbose, is_python_module=True)
    return TORCH_EXT_SO


def load_torch_binding():
    """Import the built binding (libb200vq.so first, globally, so that both share one instance)."""
    import ctypes
    import importlib.machinery
    import importlib.util
    if not os.path.exists(TORCH_EXT_SO):
        raise RuntimeError(f"{TORCH_EXT_SO} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
    ctypes.CDLL(SO_PATH, mode=ctypes.RTLD_GLOBAL)
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    loader = importlib.machinery.ExtensionFileLoader(TORCH_EXT_NAME, TORCH_EXT_SO)
    spec = importlib.util.spec_from_loader(TORCH_EXT_NAME, loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod
