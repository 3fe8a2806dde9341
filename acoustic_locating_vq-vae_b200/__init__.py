"""B200-native VectorQuantizer hot path of guy3540/Acoustic_Locating_VQ-VAE.

Only what the path needs: `csrc/` (hand-written sm_100a CUDA + the C ABI of include/b200vq.h),
the ctypes binding (`_lib`) and the host-side mirror of the reference module (`quantizer`).
The directory name carries a hyphen; import it as `b200vq` (repo-root shim) or with
`importlib.import_module("acoustic_locating_vq-vae_b200")`.
"""
from . import _lib, build as _build
from ._lib import B200VQError, load as load_library
from .quantizer import VectorQuantizer, swap_quantizers
from .onehot_linear import OneHotLinear
from .jitter import Jitter
from .time_mean import time_mean, time_mean_quantize

build_extension = _build.build
SO_PATH = _build.SO_PATH

__all__ = ["VectorQuantizer", "swap_quantizers", "OneHotLinear", "Jitter", "time_mean", "time_mean_quantize", "load_library", "build_extension", "B200VQError", "SO_PATH"]
