"""Jitter: drop-in for the reference's latent jitter.

`/root/reference/src/acoustic_locating_vq_vae/vq_vae/modules/jitter.py:47-70` walks the T time steps in Python,
draws `np.random.choice` per step and does up to T strided in-place copies on the quantizer's output.  Here the
decisions are drawn on the host with the SAME np.random calls in the same order (so a numpy seed gives the same
result, including the reference's quirk that a column is replaced with probability 1 - p, jitter.py:50), and
applied by ONE in-place gather kernel.  As in the reference the input tensor is mutated and returned, and the
replaced columns carry no gradient (their values come from a detached clone).

Attribution: the decision procedure reproduced by `draw_source_columns` (replace with a fixed probability, first / last
column take their only neighbour, otherwise left or right with equal probability) is that of the reference's jitter.py,
which carries an MIT licence header: Copyright (C) 2019 Charly Lamothe, part of VQ-VAE-Speech.  Only the sequence of
random draws is mirrored here (it is the contract a numpy seed pins); the tensor work is this repository's own kernel.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from ._lib import check


def draw_source_columns(length: int, probability: float) -> np.ndarray:
    """src[i] = original column that lands in column i; consumes np.random exactly like jitter.py:50-62."""
    src = np.arange(length, dtype=np.int32)
    for i in range(length):
        replace = [True, False][np.random.choice([1, 0], p=[probability, 1 - probability])]
        if replace:
            if i == 0:
                if length < 2:      # the reference indexes column 1 of a single-column tensor here (jitter.py:57,68)
                    raise IndexError("index 1 is out of bounds for dimension 2 with size 1")
                src[i] = i + 1
            elif i == length - 1:
                src[i] = i - 1
            else:
                src[i] = i + np.random.choice([-1, 1], p=[0.5, 0.5])
    return src


class _JitterFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, quantized, src):
        lib = _lib.load()
        B, D, T = quantized.shape
        with torch.cuda.device(quantized.device):      # the C ABI launches on the current device: follow the tensor
            check(lib.vq_jitter_apply(quantized.data_ptr(), src.data_ptr(), B * D, T, torch.cuda.current_stream(quantized.device).cuda_stream))
        ctx.mark_dirty(quantized)
        ctx.save_for_backward(src)
        return quantized

    @staticmethod
    def backward(ctx, g):
        (src,) = ctx.saved_tensors
        lib = _lib.load()
        g = g.contiguous().clone()
        B, D, T = g.shape
        with torch.cuda.device(g.device):
            check(lib.vq_jitter_backward(g.data_ptr(), src.data_ptr(), B * D, T, torch.cuda.current_stream(g.device).cuda_stream))
        return g, None


class Jitter(nn.Module):
    def __init__(self, probability=0.12):
        super().__init__()
        self._probability = probability

    def forward(self, quantized):
        if not quantized.is_cuda:
            raise RuntimeError("b200vq.Jitter runs on a B200 GPU only (no CPU fallback)")
        if quantized.dim() != 3 or not quantized.is_contiguous() or quantized.dtype != torch.float32:
            raise RuntimeError("b200vq.Jitter expects a contiguous float32 (B, D, T) tensor")
        src = torch.from_numpy(draw_source_columns(quantized.size(2), self._probability)).to(quantized.device)
        return _JitterFn.apply(quantized, src)
