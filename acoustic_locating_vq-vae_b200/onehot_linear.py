"""OneHotLinear: the consumer of the quantizer's dense one-hot as an index gather.

The reference's location head feeds `encodings.reshape(B, 201, 1024)` (train_location.py:74-75) into
`LocationModule.fc_1 = nn.Linear(201*1024, 1024)` (location_model.py:10,21): a 205 824 x 1024 GEMM on a matrix
that is 99.9 % zeros, an 843 MB weight gradient and a 211 MB one-hot that exists only to be multiplied.
`fc_1(one_hot)[b] == bias + sum_t W[:, t*K + idx[b, t]]`, so this module takes the int32 code indices
(`VectorQuantizer.last_indices.view(B, T)`) and gathers T rows of the transposed weight per sample.

Same numbers as `fc_1(flatten(one_hot))` up to fp32 summation order; the weight gradient is row-sparse
(B*T rows) and is returned as a sparse COO tensor by default (use torch.optim.SparseAdam / SGD) or densely.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from ._lib import check


class _GatherSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, idx, weight_t, bias, K, sparse_grad):
        lib = _lib.load()
        B, T = idx.shape
        O = weight_t.shape[1]
        y = torch.empty(B, O, dtype=torch.float32, device=weight_t.device)
        w = weight_t.detach()
        with torch.cuda.device(weight_t.device):       # the C ABI launches on the current device: follow the tensors
            st = torch.cuda.current_stream(weight_t.device).cuda_stream
            check(lib.vq_gather_sum_rows(idx.data_ptr(), w.data_ptr(), None if bias is None else bias.detach().data_ptr(),
                                         y.data_ptr(), B, T, K, O, st))
        ctx.save_for_backward(idx)
        ctx.shape = (weight_t.shape[0], O, K, bias is not None, sparse_grad)
        return y

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        rows_total, O, K, has_bias, sparse_grad = ctx.shape
        B, T = idx.shape
        g = g.contiguous().float()
        gw = None
        if ctx.needs_input_grad[1]:
            rows = (torch.arange(T, device=idx.device, dtype=torch.int64) * K)[None, :] + idx.long()     # (B, T)
            if sparse_grad:
                vals = g[:, None, :].expand(B, T, O).reshape(B * T, O)
                gw = torch.sparse_coo_tensor(rows.reshape(1, -1), vals, (rows_total, O))
            else:
                lib = _lib.load()
                gw = torch.zeros(rows_total, O, dtype=torch.float32, device=g.device)
                with torch.cuda.device(g.device):
                    check(lib.vq_scatter_add_rows(idx.data_ptr(), g.data_ptr(), gw.data_ptr(), B, T, K, O,
                                                  torch.cuda.current_stream(g.device).cuda_stream))
        gb = g.sum(0) if (has_bias and ctx.needs_input_grad[2]) else None
        return None, gw, gb, None, None


class OneHotLinear(nn.Module):
    """`nn.Linear(T*K, O)` applied to a flattened (B, T, K) one-hot, computed from the (B, T) code indices.

    weight_t: (T*K, O) -- the Linear's weight TRANSPOSED, so that the row of a (position, code) pair is contiguous.
    """

    def __init__(self, positions: int, num_embeddings: int, out_features: int, bias: bool = True,
                 sparse_grad: bool = True):
        super().__init__()
        if out_features % 4 != 0:
            raise ValueError("out_features must be a multiple of 4")
        self.positions, self.num_embeddings, self.out_features = positions, num_embeddings, out_features
        self.sparse_grad = sparse_grad
        lin = nn.Linear(positions * num_embeddings, out_features, bias=bias)      # same init as the reference's fc_1
        self.weight_t = nn.Parameter(lin.weight.detach().t().contiguous())
        self.bias = nn.Parameter(lin.bias.detach().clone()) if bias else None

    @classmethod
    def from_linear(cls, linear: nn.Linear, positions: int, num_embeddings: int, sparse_grad: bool = True) -> "OneHotLinear":
        assert linear.in_features == positions * num_embeddings
        m = cls.__new__(cls)
        nn.Module.__init__(m)
        m.positions, m.num_embeddings, m.out_features = positions, num_embeddings, linear.out_features
        m.sparse_grad = sparse_grad
        m.weight_t = nn.Parameter(linear.weight.detach().t().contiguous())
        m.bias = None if linear.bias is None else nn.Parameter(linear.bias.detach().clone())
        return m

    def to_linear(self) -> nn.Linear:
        lin = nn.Linear(self.positions * self.num_embeddings, self.out_features, bias=self.bias is not None,
                        device=self.weight_t.device)
        lin.weight.data.copy_(self.weight_t.detach().t())
        if self.bias is not None:
            lin.bias.data.copy_(self.bias.detach())
        return lin

    def forward(self, indices: torch.Tensor) -> torch.Tensor:
        if not indices.is_cuda:
            raise RuntimeError("b200vq.OneHotLinear runs on a B200 GPU only (no CPU fallback)")
        if indices.dim() != 2 or indices.shape[1] != self.positions:
            raise RuntimeError(f"expected (B, {self.positions}) code indices, got {tuple(indices.shape)}")
        idx = indices.to(torch.int32).contiguous()     # codes outside [0, K) contribute nothing (the kernels skip them)
        return _GatherSum.apply(idx, self.weight_t, self.bias, self.num_embeddings, self.sparse_grad)
