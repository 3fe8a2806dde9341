"""CPU oracle for the VectorQuantizer hot path -- TEST INFRASTRUCTURE ONLY.

This file restates, in plain PyTorch fp32 on the CPU, what the reference module
`src/acoustic_locating_vq_vae/vq_vae/vector_quantizer.py:29-58` computes, plus the gradients
PyTorch autograd derives from it.  It is the *checker* for the CUDA path: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import
it.  The product package never imports anything from `oracle/`.

Pinning: `check_against_reference()` imports the real reference class from `/root/reference`
(when that tree exists, i.e. in the authoring container) and asserts that this restatement is
bit-identical to it on CPU for outputs and gradients.  The golden fixtures under
`tests/golden/` were generated from the real reference class by
`tests/golden/make_golden.py`; `tests/test_oracle.py` checks this file against them.

Two restatements live here:

* `forward_dense` / `forward_backward_dense` follow the reference op by op (same aten calls, so
  the same MKL / oneDNN kernels and the same rounding) -- this is what is timed as the CPU
  baseline, because it does exactly the work the reference does.
* `forward_compact` is the algebraically reduced form (index gather instead of one-hot GEMMs,
  bincount instead of a dense mean) used to state the identities the CUDA kernels rely on
  (SURVEY.md Appendix A.3).
"""
from __future__ import annotations

import os
import sys
from dataclasses import dataclass
from typing import Optional

import torch

REFERENCE_ROOT = "/root/reference"


@dataclass
class VQResult:
    loss: torch.Tensor          # 0-dim fp32
    quantized: torch.Tensor     # inputs.shape fp32 (straight-through value)
    perplexity: torch.Tensor    # 0-dim fp32
    encodings: torch.Tensor     # (N, K) fp32 one-hot
    indices: torch.Tensor       # (N,) int64
    dz: Optional[torch.Tensor] = None   # d(objective)/d(inputs)
    dE: Optional[torch.Tensor] = None   # d(objective)/d(codebook) or None when frozen


def init_codebook(num_embeddings: int, embedding_dim: int, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """Codebook initialisation of the reference: U(-1/K, 1/K) (vector_quantizer.py:15-16).

    (The reference first draws nn.Embedding's N(0,1) init and then overwrites it; only the
    uniform draw survives, but both consume the global RNG -- see `make_golden.py`.)
    """
    w = torch.empty(num_embeddings, embedding_dim, dtype=torch.float32)
    w.uniform_(-1.0 / num_embeddings, 1.0 / num_embeddings, generator=generator)
    return w


def flatten_rows(inputs: torch.Tensor, embedding_dim: int) -> torch.Tensor:
    """vector_quantizer.py:30-32 -- `inputs.view(-1, D)` with NO permute.

    Rows are consecutive D-float chunks of the contiguous buffer.  `.view` raises RuntimeError
    for a non-contiguous input or when numel is not divisible by D; so does this.
    """
    return inputs.view(-1, embedding_dim)


def distances(flat: torch.Tensor, codebook: torch.Tensor) -> torch.Tensor:
    """vector_quantizer.py:34-36 -- ((|z|^2 + |E|^2) - 2 z E^T), in that evaluation order."""
    z_sq = torch.sum(flat ** 2, dim=1, keepdim=True)
    e_sq = torch.sum(codebook ** 2, dim=1)
    return z_sq + e_sq - 2 * torch.matmul(flat, codebook.t())


def forward_dense(inputs: torch.Tensor, codebook: torch.Tensor, commitment_cost: float,
                  train_vq: bool = True) -> VQResult:
    """Op-by-op restatement of vector_quantizer.py:29-58 (autograd-capable)."""
    K, D = codebook.shape
    shape = inputs.shape
    flat = flatten_rows(inputs, D)
    dist = distances(flat, codebook)
    # :38 first minimal index on ties (torch.argmin semantics)
    idx = torch.argmin(dist, dim=1).unsqueeze(1)
    # :39-40 dense one-hot, fp32, on inputs.device
    onehot = torch.zeros(idx.shape[0], K, device=inputs.device)
    onehot.scatter_(1, idx, 1)
    # :43 gather written as a dense GEMM
    q = torch.matmul(onehot, codebook).view(shape)
    # :46-50 commitment + codebook terms
    mse = torch.nn.functional.mse_loss
    e_latent = mse(q.detach(), inputs, reduction="mean")
    if train_vq:
        q_latent = mse(q, inputs.detach(), reduction="mean")
    else:
        q_latent = mse(q.detach(), inputs.detach(), reduction="mean")
    loss = q_latent + commitment_cost * e_latent                     # :52
    q_st = inputs + (q - inputs).detach()                            # :54 straight-through
    probs = torch.mean(onehot, dim=0)                                # :55
    perplexity = torch.exp(-torch.sum(probs * torch.log(probs + 1e-10)))   # :56
    return VQResult(loss, q_st.contiguous(), perplexity, onehot, idx.squeeze(1))


def forward_backward_dense(inputs: torch.Tensor, codebook: torch.Tensor, commitment_cost: float,
                           train_vq: bool = True, g_quantized: Optional[torch.Tensor] = None,
                           g_loss: float = 1.0) -> VQResult:
    """Forward + autograd backward of objective `g_loss*loss + sum(g_quantized*quantized)`.

    With g_quantized = ones and g_loss = 1 this is `(loss + quantized.sum()).backward()`, the
    workload SURVEY.md section 8(d) defines for the throughput metric.
    """
    z = inputs.detach().clone().requires_grad_(True)
    E = codebook.detach().clone().requires_grad_(True)
    res = forward_dense(z, E, commitment_cost, train_vq)
    if g_quantized is None:
        g_quantized = torch.ones_like(res.quantized)
    objective = g_loss * res.loss + (g_quantized * res.quantized).sum()
    objective.backward()
    res.dz = z.grad.detach()
    res.dE = None if E.grad is None else E.grad.detach()
    res.loss = res.loss.detach()
    res.quantized = res.quantized.detach()
    return res


def forward_compact(inputs: torch.Tensor, codebook: torch.Tensor, commitment_cost: float,
                    indices: Optional[torch.Tensor] = None) -> VQResult:
    """The reduced form the CUDA kernels implement (no N x K temporaries after the argmin).

    quantized = z + (E[idx] - z)            (vector_quantizer.py:43,54)
    loss      = m + beta*m, m = mean((E[idx]-z)^2)   (:46-52)
    perplexity= exp(-sum c/N * log(c/N + 1e-10)), c = bincount(idx)   (:55-56)
    """
    K, D = codebook.shape
    flat = flatten_rows(inputs, D)
    if indices is None:
        indices = torch.argmin(distances(flat, codebook), dim=1)
    q = codebook[indices]
    diff = q - flat
    m = torch.mean(diff * diff)
    loss = m + commitment_cost * m
    counts = torch.bincount(indices, minlength=K).to(torch.float32)
    probs = counts / flat.shape[0]
    perplexity = torch.exp(-torch.sum(probs * torch.log(probs + 1e-10)))
    onehot = torch.nn.functional.one_hot(indices, K).to(torch.float32)
    return VQResult(loss, (flat + diff).view(inputs.shape), perplexity, onehot, indices)


def backward_compact(inputs: torch.Tensor, codebook: torch.Tensor, indices: torch.Tensor,
                     commitment_cost: float, g_quantized: Optional[torch.Tensor], g_loss: float,
                     train_vq: bool = True):
    """Closed-form gradients of vector_quantizer.py:46-54.

    dz = g_q + g_loss * beta * 2 (z - q) / (N D)
    dE = index_add(idx, g_loss * 2 (q - z) / (N D))   or None when the codebook is frozen.
    The straight-through output carries NO gradient to the codebook.
    """
    K, D = codebook.shape
    flat = flatten_rows(inputs, D)
    q = codebook[indices]
    numel = flat.numel()
    dz = (g_loss * commitment_cost * 2.0 / numel) * (flat - q)
    if g_quantized is not None:
        dz = dz + g_quantized.reshape(flat.shape)
    dE = None
    if train_vq:
        dE = torch.zeros_like(codebook)
        dE.index_add_(0, indices, (g_loss * 2.0 / numel) * (q - flat))
    return dz.view(inputs.shape), dE


# --------------------------------------------------------------------------------------------
# Pinning against the real reference (authoring container only; /root/reference is absent on
# the GPU box, and nothing run there calls this).
# --------------------------------------------------------------------------------------------

def import_reference_class():
    """Import the unmodified reference VectorQuantizer; None when /root/reference is absent."""
    if not os.path.isdir(REFERENCE_ROOT):
        return None
    for p in (REFERENCE_ROOT, os.path.join(REFERENCE_ROOT, "src")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from acoustic_locating_vq_vae.vq_vae.vector_quantizer import VectorQuantizer  # type: ignore
    return VectorQuantizer


def check_against_reference(shapes=((2, 4, 3, 8), (8, 64, 201, 1024), (4, 128, 500, 1024)),
                            verbose: bool = False) -> bool:
    """Assert forward_dense/forward_backward_dense == reference class, bit for bit, on CPU.

    shapes: tuples (B, D, T, K).  Returns False (and checks nothing) without /root/reference.
    """
    Ref = import_reference_class()
    if Ref is None:
        return False
    for (B, D, T, K) in shapes:
        for init in ("uniform", "normal"):
            for train_vq in (True, False):
                torch.manual_seed(7)
                ref = Ref(K, D, 0.25)
                if init == "normal":
                    ref._embedding.weight.data.normal_()
                ref.set_train_vq(train_vq)
                z = torch.randn(B, D, T, requires_grad=True)
                g = torch.randn(B, D, T)
                loss, q, perp, enc = ref(z)
                (0.7 * loss + (g * q).sum()).backward()
                mine = forward_backward_dense(z.detach(), ref._embedding.weight.detach(), 0.25,
                                              train_vq=train_vq, g_quantized=g, g_loss=0.7)
                assert torch.equal(mine.loss, loss.detach()), (B, D, T, K, init, "loss")
                assert torch.equal(mine.quantized, q.detach()), (B, D, T, K, init, "quantized")
                assert torch.equal(mine.perplexity, perp.detach()), (B, D, T, K, init, "perplexity")
                assert torch.equal(mine.encodings, enc), (B, D, T, K, init, "encodings")
                assert torch.equal(mine.dz, z.grad), (B, D, T, K, init, "dz")
                if train_vq:
                    assert torch.equal(mine.dE, ref._embedding.weight.grad), (B, D, T, K, init, "dE")
                else:
                    assert ref._embedding.weight.grad is None and mine.dE is None
                if verbose:
                    print(f"oracle == reference  B={B} D={D} T={T} K={K} init={init} train_vq={train_vq}")
    return True


if __name__ == "__main__":
    ok = check_against_reference(verbose=True)
    print("pinned against /root/reference" if ok else "reference tree absent: nothing checked")
