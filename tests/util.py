"""Helpers shared by CPU and GPU parity tests."""
from __future__ import annotations

import hashlib

import numpy as np

# A row is an fp32 near-tie when, in float64, the distance of the code we picked and of the code
# the reference picked differ by less than `tol * (|z|^2 + |E|^2)` -- i.e. by a few fp32 ulps of the
# distance magnitude, below what any fp32 GEMM accumulation order can resolve.
NEAR_TIE_EXACT = 2.0 ** -21      # CUDA-core fp32 FMA-chain path vs the reference's MKL sgemm
NEAR_TIE_TENSOR = 2.0 ** -19     # 3xTF32 tensor path (dropped lo*lo term + tensor-core accumulation)


def sha16(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def near_tie_report(z_rows: np.ndarray, E: np.ndarray, idx_mine: np.ndarray, idx_ref: np.ndarray):
    """Returns (n_mismatch, worst relative fp64 gap over the mismatching rows)."""
    mism = np.nonzero(idx_mine != idx_ref)[0]
    if mism.size == 0:
        return 0, 0.0
    zr = z_rows[mism].astype(np.float64)
    e_ref = E[idx_ref[mism]].astype(np.float64)
    e_me = E[idx_mine[mism]].astype(np.float64)
    d_ref = ((zr - e_ref) ** 2).sum(1)
    d_me = ((zr - e_me) ** 2).sum(1)
    scale = (zr ** 2).sum(1) + np.maximum((e_ref ** 2).sum(1), (e_me ** 2).sum(1))
    return int(mism.size), float((np.abs(d_ref - d_me) / scale).max())


def rel_err(a, b) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(a - b).max() / denom)
