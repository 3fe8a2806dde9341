"""ctypes front-end of oracle/vq_oracle.c -- TEST INFRASTRUCTURE ONLY (see that file's header).

The C oracle fixes the fp32 summation order (sequential FMA chains) that the reference leaves to
its BLAS (vector_quantizer.py:34-36), so the CUDA "exact" path can be compared bit for bit.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libvq_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "vq_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int32)
        dp = ctypes.POINTER(ctypes.c_double)
        i64, i32, f32 = ctypes.c_int64, ctypes.c_int, ctypes.c_float
        L.vq_oracle_abi_version.restype = ctypes.c_int
        L.vq_oracle_num_threads.restype = ctypes.c_int
        L.vq_oracle_code_norms.argtypes = [fp, i32, i32, fp]
        L.vq_oracle_argmin.argtypes = [fp, fp, i64, i32, i32, ip, fp]
        L.vq_oracle_quantize.argtypes = [fp, fp, ip, i64, i32, i32, f32, fp, fp, fp, fp, fp, dp]
        L.vq_oracle_backward.argtypes = [fp, f32, fp, fp, ip, i64, i64, i64, i32, i32, f32, fp, fp]
        L.vq_oracle_step.argtypes = [fp, fp, fp, f32, i64, i32, i32, f32, i32, ip, fp, fp, fp, fp, fp, fp]
        _lib = L
    return _lib


def _f(a):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _i(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def num_threads() -> int:
    return int(lib().vq_oracle_num_threads())


def code_norms(E):
    E = _c(E)
    K, D = E.shape
    out = np.empty(K, np.float32)
    lib().vq_oracle_code_norms(_f(E), K, D, _f(out))
    return out


def argmin(z_rows, E, want_dist: bool = False):
    """z_rows: (N, D) fp32 rows of the flattened input; returns int32 indices (and distances)."""
    z = _c(z_rows)
    E = _c(E)
    N, D = z.shape
    K = E.shape[0]
    idx = np.empty(N, np.int32)
    dist = np.empty((N, K), np.float32) if want_dist else None
    lib().vq_oracle_argmin(_f(z), _f(E), N, K, D, _i(idx), _f(dist))
    return (idx, dist) if want_dist else idx


def quantize(z_rows, E, idx, beta: float, want_onehot: bool = False):
    z = _c(z_rows)
    E = _c(E)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    N, D = z.shape
    K = E.shape[0]
    q = np.empty_like(z)
    onehot = np.empty((N, K), np.float32) if want_onehot else None
    hist = np.empty(K, np.float32)
    loss = np.zeros(1, np.float32)
    perp = np.zeros(1, np.float32)
    sse = ctypes.c_double(0.0)
    lib().vq_oracle_quantize(_f(z), _f(E), _i(idx), N, K, D, beta, _f(q), _f(onehot), _f(hist),
                             _f(loss), _f(perp), ctypes.byref(sse))
    return dict(quantized=q, onehot=onehot, hist=hist, loss=loss[0], perplexity=perp[0], sse=sse.value)


def backward(g_q, g_loss: float, z_rows, E, idx, beta: float, train_vq: bool = True,
             n_rows_dz=None, n_rows_dE=None):
    z = _c(z_rows)
    E = _c(E)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    g = None if g_q is None else _c(g_q).reshape(z.shape)
    N, D = z.shape
    K = E.shape[0]
    dz = np.empty_like(z)
    dE = np.empty_like(E) if train_vq else None
    lib().vq_oracle_backward(_f(g), g_loss, _f(z), _f(E), _i(idx), N,
                             N if n_rows_dz is None else n_rows_dz,
                             N if n_rows_dE is None else n_rows_dE, K, D, beta, _f(dz), _f(dE))
    return dz, dE


def step(z_rows, E, g_q, g_loss: float, beta: float, train_vq: bool = True):
    """Whole forward+backward on the CPU (the 'port' baseline bench.py can time)."""
    z = _c(z_rows)
    E = _c(E)
    g = None if g_q is None else _c(g_q).reshape(z.shape)
    N, D = z.shape
    K = E.shape[0]
    idx = np.empty(N, np.int32)
    q = np.empty_like(z)
    hist = np.empty(K, np.float32)
    loss = np.zeros(1, np.float32)
    perp = np.zeros(1, np.float32)
    dz = np.empty_like(z)
    dE = np.empty_like(E)
    lib().vq_oracle_step(_f(z), _f(E), _f(g), g_loss, N, K, D, beta, int(train_vq), _i(idx), _f(q),
                         _f(hist), _f(loss), _f(perp), _f(dz), _f(dE))
    return dict(indices=idx, quantized=q, hist=hist, loss=loss[0], perplexity=perp[0], dz=dz,
                dE=dE if train_vq else None)
