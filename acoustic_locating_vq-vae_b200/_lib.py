"""ctypes binding of include/b200vq.h (libb200vq.so).  Fails loudly when the library is missing.

This is the whole FFI surface: plain pointers, sizes and a cudaStream_t.  Nothing here computes.
"""
from __future__ import annotations

import ctypes
import os

from .build import SO_PATH

ABI_VERSION = 18

FLAG_ONEHOT = 1 << 0
FLAG_TRAIN_VQ = 1 << 1
FLAG_EXACT = 1 << 2
FLAG_DEFER_STATS = 1 << 3
FLAG_NO_QUANT = 1 << 4
FLAG_ZERO_DE = 1 << 5
FLAG_TC_1CTA = 1 << 6
FLAG_NO_FUSE = 1 << 7
FLAG_STATE_READY = 1 << 8
FLAG_NO_SCREEN = 1 << 9
FLAG_SCREEN = 1 << 10
FLAG_BWD_FLAT = 1 << 12
FLAG_BWD_PRIVATE = 1 << 14

_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_int = ctypes.c_int
_f32 = ctypes.c_float
_sz = ctypes.c_size_t

# name -> (restype, argtypes); every symbol include/b200vq.h declares
SIGNATURES = {
    "vq_abi_version": (_int, []),
    "vq_last_error": (ctypes.c_char_p, []),
    "vq_device_check": (_int, []),
    "vq_forward_uses_tensor_path": (_int, [_i64, _int, _int, _int]),
    "vq_launch_count": (_i64, []),
    "vq_profile_enable": (None, [_int]),
    "vq_profile_read": (_int, [_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_i64)]),
    "vq_debug_set_trace": (None, [_vp]),
    "vq_prepare_codebook": (_int, [_vp, _int, _int, _vp, _vp, _vp, _vp]),
    "vq_prepare_step": (_int, [_vp, _int, _int, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "vq_workspace_bytes": (_sz, [_i64, _int, _int, _int]),
    "vq_forward": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _int, _int, _f32, _int,
                          _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "vq_step_forward": (_int, [_vp, _vp, _i64, _int, _int, _f32, _int, _vp, _vp, _vp, _vp,
                               _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "vq_finalize_stats": (_int, [_vp, _vp, _i64, _int, _int, _f32, _vp, _vp, _vp]),
    "vq_onehot": (_int, [_vp, _i64, _int, _vp, _vp]),
    "vq_backward": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _int, _int, _f32, _int, _vp, _vp, _vp]),
    "vq_step_backward": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _int, _int, _f32, _int, _vp, _vp, _vp, _sz, _int, _vp]),
    "vq_backward_path": (_int, [_i64, _int, _int, _int]),
    "vq_gather_sum_rows": (_int, [_vp, _vp, _vp, _vp, _int, _int, _int, _int, _vp]),
    "vq_scatter_add_rows": (_int, [_vp, _vp, _vp, _int, _int, _int, _int, _vp]),
    "vq_time_mean": (_int, [_vp, _i64, _int, _vp, _vp]),
    "vq_time_mean_backward": (_int, [_vp, _i64, _int, _vp, _vp]),
    "vq_jitter_apply": (_int, [_vp, _vp, _i64, _int, _vp]),
    "vq_jitter_backward": (_int, [_vp, _vp, _i64, _int, _vp]),
    "vq_dp_recv_lines": (_i64, [_int, _i64]),
    "vq_dp_create": (_int, [ctypes.POINTER(_vp), ctypes.POINTER(_vp), _vp, _vp, _int, _int, _i64, ctypes.c_uint32, ctypes.POINTER(_vp)]),
    "vq_dp_destroy": (None, [_vp]),
    "vq_dp_allreduce": (_int, [_vp, _vp, _vp, _vp]),
    "vq_step_backward_dp": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _int, _int, _f32, _int, _vp, _vp, _vp, _vp, _vp, _sz, _int, _vp]),
    "vq_dp_allreduce_start": (_int, [_vp, _vp, _vp, _vp]),
    "vq_dp_wait": (_int, [_vp, _int, _vp]),
    "vq_dp_status": (_int, [_vp, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32), _vp]),
    "vq_dp_emulate": (_int, [_int, _int, _i64, ctypes.POINTER(_vp), ctypes.POINTER(_vp), _int, ctypes.c_uint32,
                             ctypes.POINTER(ctypes.c_uint32), _vp]),
    "vq_host_ctx_create": (_int, [_i64, _int, _int, ctypes.POINTER(_vp)]),
    "vq_host_ctx_destroy": (None, [_vp]),
    "vq_host_set_codebook": (_int, [_vp, _vp]),
    "vq_host_step_async": (_int, [_vp, _int, _vp, _vp, _i64, _i64, _f32, _int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "vq_host_timer_start": (_int, [_vp]),
    "vq_host_timer_stop_ms": (_int, [_vp, ctypes.POINTER(_f32)]),
    "vq_host_wait": (_int, [_vp, _int]),
    "vq_host_lane_buffers": (_int, [_vp, _int, ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.POINTER(_vp)]),
}

_lib = None


class B200VQError(RuntimeError):
    pass


def load():
    """dlopen csrc/libb200vq.so (built by build.py / __graft_entry__.build()); no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise B200VQError(
            f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "The B200 VectorQuantizer has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(SO_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    got = lib.vq_abi_version()
    if got != ABI_VERSION:
        raise B200VQError(f"libb200vq.so ABI {got} != binding ABI {ABI_VERSION}: rebuild the extension")
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().vq_last_error()
        raise B200VQError(f"b200vq error {rc}: {msg.decode() if msg else '?'}")
