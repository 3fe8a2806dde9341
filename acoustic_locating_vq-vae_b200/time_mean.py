"""TimeMean: the pooling step the reference may put between `_pre_vq_conv` and the quantizer.

`ConvolutionalVQVAE.forward` with `encoder_average_pooling=True` does `z = torch.mean(z, dim=2, keepdim=True)`
(/root/reference/src/acoustic_locating_vq_vae/vq_vae/convolutional_vq_vae.py:96-97) before `self._vq(z)`; the quantizer
then sees one row per batch item.  Here that mean is one warp-per-row CUDA kernel (vq_time_mean) launched with the
programmatic-launch attribute, so the quantizer's prepare + forward launches chain behind it without a gap, and its
backward (dz / T broadcast over time) is one kernel as well.  `time_mean_quantize(vq, x)` is the fused call:
`vq(mean(x, 2, keepdim=True))` with the reference's 4-tuple.  The mean is summed in a fixed order that differs from
torch.mean's by fp32 rounding (<= a few 1e-7 relative); given the pooled z, indices are bit-exact vs the oracle as always.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check


class _TimeMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        lib = _lib.load()
        B, D, T = x.shape
        z = torch.empty(B, D, 1, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            check(lib.vq_time_mean(x.data_ptr(), B * D, T, z.data_ptr(), torch.cuda.current_stream(x.device).cuda_stream))
        ctx.shape = (B, D, T)
        return z

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        B, D, T = ctx.shape
        g = g.contiguous().float()
        dx = torch.empty(B, D, T, dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            check(lib.vq_time_mean_backward(g.data_ptr(), B * D, T, dx.data_ptr(), torch.cuda.current_stream(g.device).cuda_stream))
        return dx


def time_mean(x: torch.Tensor) -> torch.Tensor:
    """(B, D, T) -> (B, D, 1): `torch.mean(x, dim=2, keepdim=True)` as one CUDA kernel (no CPU fallback)."""
    if not x.is_cuda:
        raise RuntimeError("b200vq.time_mean runs on a B200 GPU only (no CPU fallback)")
    if x.dim() != 3 or x.dtype != torch.float32 or not x.is_contiguous():
        raise RuntimeError("b200vq.time_mean expects a contiguous float32 (B, D, T) tensor")
    return _TimeMean.apply(x)


def time_mean_quantize(vq, x: torch.Tensor):
    """`vq(torch.mean(x, dim=2, keepdim=True))` (convolutional_vq_vae.py:96-98): (loss, quantized (B, D, 1), perplexity, encodings)."""
    return vq(time_mean(x))
