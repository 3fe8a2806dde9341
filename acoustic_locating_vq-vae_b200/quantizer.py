"""VectorQuantizer: host-side mirror of the reference module, backed by libb200vq.so.

Mirrors `/root/reference/src/acoustic_locating_vq_vae/vq_vae/vector_quantizer.py:8-58`:
same constructor, same attributes (`_embedding`, `_embedding_dim`, `_num_embeddings`,
`_commitment_cost`, `_train_vq`, `_flag_flatten`), same state-dict key (`_embedding.weight`),
same `forward(inputs) -> (loss, quantized, perplexity, encodings)`, same error behaviour for
inputs that `.view(-1, D)` rejects.  All arithmetic happens in hand-written sm_100a CUDA behind
the C ABI in include/b200vq.h; PyTorch only owns the buffers and the stream.  There is no CPU
path: a non-CUDA input raises.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

import os

from . import _lib, build as _build, parallel
from ._lib import (FLAG_EXACT, FLAG_ONEHOT, FLAG_TRAIN_VQ, FLAG_ZERO_DE, check)


_ext = None


def _binding():
    """The C++ autograd node (csrc/torch_binding.cpp): same logic as _VQFunction below, without ~120 us of Python per step.
    B200VQ_PY_AUTOGRAD=1 selects the Python node (debugging); both drive the same kernels through the same C ABI."""
    global _ext
    if _ext is None:
        _lib.load()
        _ext = False if os.environ.get("B200VQ_PY_AUTOGRAD") == "1" else _build.load_torch_binding()
    return _ext


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream(dev: Optional[torch.device] = None) -> int:
    """Raw handle of the current CUDA stream (the fast accessor when this torch has it: this sits on the eager path)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device() if dev is None or dev.index is None else dev.index)
    return torch.cuda.current_stream(dev).cuda_stream


class _Buffers:
    """Per-module device scratch, grown on demand and reused (stream-ordered, one stream per module)."""

    def __init__(self):
        self.ws = None
        self.code = None     # (e_norm2, E_hi, E_lo)
        self.ws_bytes = {}   # (N, K, D) -> vq_workspace_bytes

    def workspace(self, nbytes: int, device) -> torch.Tensor:
        if self.ws is None or self.ws.numel() < nbytes or self.ws.device != device:
            self.ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        return self.ws

    def codebook(self, K: int, D: int, device):
        c = self.code
        if c is None or c[1].shape != (K, D) or c[1].device != device:
            self.code = (torch.empty(K, dtype=torch.float32, device=device),
                         torch.empty(K, D, dtype=torch.float32, device=device),
                         torch.empty(K, D, dtype=torch.float32, device=device))
        return self.code


class _VQFunction(torch.autograd.Function):
    """forward = vq_step_forward (codebook preparation + fused forward behind one C call); backward = vq_backward,
    followed under data parallelism by ONE sum all-reduce of the packed [dE | usage histogram | squared error]."""

    @staticmethod
    def forward(ctx, inputs, weight, module, flags, want_onehot, pack):
        lib = _lib.load()
        K, D = weight.shape
        flat = inputs.view(-1, D)                     # vector_quantizer.py:32 (raises like the reference)
        N = flat.shape[0]
        dev = inputs.device
        st = _stream(dev)
        w = weight.detach()
        if not w.is_contiguous():
            w = w.contiguous()
        bufs: _Buffers = module._bufs
        e_norm2, e_hi, e_lo = bufs.codebook(K, D, dev)
        q_out = torch.empty_like(inputs)
        idx = torch.empty(N, dtype=torch.int32, device=dev)
        packed = None
        if pack:
            # data parallel: the forward writes hist | sse straight behind the slot of dE in the packed step buffer
            packed = torch.empty(K * D + K + 3, dtype=torch.float32, device=dev)
            scal = packed[K * D:]
        else:
            scal = torch.empty(K + 3, dtype=torch.float32, device=dev)     # [hist (K) | sse | loss | perplexity]
        onehot = torch.empty(N, K, dtype=torch.float32, device=dev) if want_onehot else None
        fl = flags | (FLAG_ONEHOT if want_onehot else 0)
        key = (N, K, D)
        nbytes = bufs.ws_bytes.get(key)
        if nbytes is None:
            nbytes = bufs.ws_bytes[key] = lib.vq_workspace_bytes(N, K, D, 0)
        ws = bufs.workspace(nbytes, dev)
        sp = scal.data_ptr()
        check(lib.vq_step_forward(_ptr(flat), _ptr(w), N, K, D, float(module._commitment_cost), fl,
                                  _ptr(e_norm2), _ptr(e_hi), _ptr(e_lo), None, _ptr(q_out), _ptr(idx), _ptr(onehot),
                                  sp, sp + 4 * K, sp + 4 * (K + 1), sp + 4 * (K + 2), _ptr(ws), ws.numel(), st))
        loss = scal[K + 1]
        perplexity = scal[K + 2]
        ctx.save_for_backward(inputs, weight, idx)
        ctx.module = module
        ctx.packed = packed
        ctx.train_vq = bool(module._train_vq)
        if onehot is not None:
            ctx.mark_non_differentiable(perplexity, idx, onehot)
        else:
            ctx.mark_non_differentiable(perplexity, idx)
        ctx.set_materialize_grads(False)            # no zero tensors for the outputs that carry no gradient (the one-hot: 211 MB)
        module.__dict__["_last_stats"] = scal       # plain attribute: skip nn.Module.__setattr__ on the eager path
        return loss, q_out, perplexity, onehot, idx

    @staticmethod
    def backward(ctx, g_loss, g_q, _g_perp, _g_onehot, _g_idx):
        lib = _lib.load()
        inputs, weight, idx = ctx.saved_tensors
        module = ctx.module
        K, D = weight.shape
        N = idx.shape[0]
        dev = inputs.device
        st = _stream(dev)
        need_dz = ctx.needs_input_grad[0]
        need_dE = ctx.needs_input_grad[1] and ctx.train_vq
        if g_loss is None:
            g_loss = torch.zeros((), dtype=torch.float32, device=dev)
        elif g_loss.dtype != torch.float32:
            g_loss = g_loss.float()
        if g_q is not None:
            g_q = g_q.contiguous()
            if g_q.dtype != torch.float32:
                g_q = g_q.float()
        pg = module.process_group
        world = 1
        if pg is not None or module.data_parallel:
            import torch.distributed as dist
            world = dist.get_world_size(pg)
        dz = torch.empty_like(inputs) if need_dz else None
        w = weight.detach()
        if not w.is_contiguous():
            w = w.contiguous()
        beta = float(module._commitment_cost)
        n_dz, n_dE = max(N, 1), max(N, 1) * world
        if not need_dE:
            if need_dz:
                check(lib.vq_backward(_ptr(g_q), _ptr(g_loss), _ptr(inputs), _ptr(w), _ptr(idx), N, n_dz, n_dE, K, D, beta, 0,
                                      _ptr(dz), None, st))
            return dz, None, None, None, None, None
        packed = ctx.packed
        if world == 1 or packed is None:
            dE = torch.empty(K, D, dtype=torch.float32, device=dev)
            check(lib.vq_backward(_ptr(g_q), _ptr(g_loss), _ptr(inputs), _ptr(w), _ptr(idx), N, n_dz, n_dE, K, D, beta,
                                  FLAG_TRAIN_VQ | FLAG_ZERO_DE, _ptr(dz), _ptr(dE), st))
            if world > 1:       # the codebook was frozen when the forward ran: nothing was packed -- plain all-reduce of dE
                import torch.distributed as dist
                dist.all_reduce(dE, group=pg)
            return dz, dE, None, None, None, None
        # data parallel: dE lands in front of the statistics the forward left in the packed buffer; ONE all-reduce
        n = K * D + K + 1
        ex = module._peer_exchange(K, D, dev, pg)
        check(lib.vq_backward(_ptr(g_q), _ptr(g_loss), _ptr(inputs), _ptr(w), _ptr(idx), N, n_dz, n_dE, K, D, beta,
                              FLAG_TRAIN_VQ | FLAG_ZERO_DE, _ptr(dz), _ptr(packed), st))
        if ex is not None:
            reduced = torch.empty(n, dtype=torch.float32, device=dev)      # fresh: autograd may keep dE as the gradient
            ex.allreduce(packed[:n], reduced, st)
        else:
            reduced = parallel.all_reduce_packed(packed[:n], pg)
        module.__dict__["_global_stats"] = (reduced[K * D:], N * world)
        return dz, reduced[:K * D].view(K, D), None, None, None, None


class VectorQuantizer(nn.Module):
    """Drop-in for the reference `VectorQuantizer` (vector_quantizer.py:8-58) on B200.

    Extra keyword-only options (defaults reproduce the reference exactly):
      return_encodings  False skips the dense (N, K) one-hot (4 K bytes per row of HBM traffic);
                        the 4th return value is then None.  `last_indices` always holds the codes.
      exact             True computes every distance on CUDA cores in the oracle's fp32 FMA-chain order; False
                        (default) uses the tcgen05 path whenever the shape allows: one TF32 screening pass plus an
                        exact fp32 refine of the candidates (K % 256 == 0, D in {32,64,96,128,192,256}; indices
                        bit-exact vs oracle/vq_oracle.c), else 3xTF32 (K % 128 == 0, D in {32,64,96,128}).
      process_group     set (or `data_parallel=True` for the default group) to all-reduce the packed
                        [dE | usage histogram | squared error] once per step across data-parallel ranks.
    """

    def __init__(self, num_embeddings, embedding_dim, commitment_cost, flag_flatten=True, *,
                 return_encodings: bool = True, exact: bool = False, process_group=None,
                 data_parallel: bool = False):
        super().__init__()
        self._embedding_dim = embedding_dim
        self._num_embeddings = num_embeddings
        # vector_quantizer.py:15-16: nn.Embedding then U(-1/K, 1/K)
        self._embedding = nn.Embedding(self._num_embeddings, self._embedding_dim)
        self._embedding.weight.data.uniform_(-1 / self._num_embeddings, 1 / self._num_embeddings)
        self._commitment_cost = commitment_cost
        self._train_vq = True
        self._flag_flatten = flag_flatten
        self.return_encodings = return_encodings
        self.exact = exact
        self.process_group = process_group
        self.data_parallel = data_parallel
        self._bufs = _Buffers()
        self._last_stats = None
        self._global_stats = None
        self._peer_ex = None
        self.last_indices = None

    # -- reference accessors (vector_quantizer.py:23-27) ------------------------------------------
    def get_embedding_dim(self):
        return self._embedding_dim

    def set_train_vq(self, train_vq):
        self._train_vq = train_vq

    # -- pickling: scratch buffers and process groups do not travel -------------------------------
    def __getstate__(self):
        s = dict(self.__dict__)
        s["_bufs"] = None
        s["_last_stats"] = None
        s["_global_stats"] = None
        s["last_indices"] = None
        s["process_group"] = None
        s["_peer_ex"] = None
        return s

    def __setstate__(self, s):
        self.__dict__.update(s)
        self._bufs = _Buffers()

    def forward(self, inputs):
        if not isinstance(inputs, torch.Tensor):
            raise TypeError("VectorQuantizer expects a tensor")
        if not inputs.is_cuda:
            raise RuntimeError("b200vq.VectorQuantizer runs on a B200 GPU only (no CPU fallback): "
                               f"got a tensor on {inputs.device}")
        if inputs.dtype != torch.float32:
            raise RuntimeError(f"b200vq.VectorQuantizer expects float32 inputs, got {inputs.dtype}")
        weight = self._embedding.weight
        if weight.device != inputs.device:
            raise RuntimeError(f"codebook on {weight.device} but inputs on {inputs.device}")
        # vector_quantizer.py:32: `.view` raises RuntimeError for incompatible strides / sizes
        flat = inputs.view(-1, self._embedding_dim)
        if not flat.is_contiguous():
            raise RuntimeError("view size is not compatible with input tensor's size and stride "
                               "(b200vq needs a contiguous input, like the reference's .view)")
        flags = FLAG_EXACT if self.exact else 0
        # data parallel with a training codebook: the forward writes its statistics straight into the packed step buffer
        # (decided here: grad mode is off inside autograd.Function.forward)
        dp = self.process_group is not None or self.data_parallel
        pack = bool(dp and self._train_vq and weight.requires_grad and torch.is_grad_enabled())
        ext = _binding()
        if ext:
            K, D = self._num_embeddings, self._embedding_dim
            dev = inputs.device
            N = inputs.numel() // D
            world, dp_ptr = 1, 0
            if dp:
                import torch.distributed as dist
                world = dist.get_world_size(self.process_group)
                if pack:
                    ex = self._peer_exchange(K, D, dev, self.process_group)
                    dp_ptr = 0 if ex is None else int(ex.ctx.value)
            if not (pack and dp_ptr == 0):     # data parallel over NCCL (no NVLink exchange): the Python node below handles it
                bufs = self._bufs
                e_norm2, e_hi, e_lo = bufs.codebook(K, D, dev)
                nbytes = bufs.ws_bytes.get((N, K, D))
                if nbytes is None:
                    nbytes = bufs.ws_bytes[(N, K, D)] = ext.workspace_bytes(N, K, D)
                ws = bufs.workspace(nbytes, dev)
                loss, quantized, perplexity, onehot, idx, stats, reduced = ext.vq_apply(
                    inputs, weight, float(self._commitment_cost), flags, bool(self.return_encodings), bool(self._train_vq), world, dp_ptr, pack,
                    e_norm2, e_hi, e_lo, ws)
                d = self.__dict__
                d["last_indices"] = idx
                d["_last_stats"] = stats[K * D:] if pack else stats
                if pack:
                    d["_global_stats"] = (reduced[K * D:], N * world)      # filled by the backward's all-reduce
                return loss, quantized, perplexity, (onehot if self.return_encodings else None)
        if inputs.device.index != torch.cuda.current_device():      # the C ABI launches on the current device
            with torch.cuda.device(inputs.device):
                loss, quantized, perplexity, encodings, idx = _VQFunction.apply(
                    inputs, weight, self, flags, bool(self.return_encodings), pack)
        else:
            loss, quantized, perplexity, encodings, idx = _VQFunction.apply(
                inputs, weight, self, flags, bool(self.return_encodings), pack)
        self.__dict__["last_indices"] = idx
        return loss, quantized, perplexity, encodings

    def _peer_exchange(self, K, D, dev, pg):
        """Lazily set up the NVLink exchange (parallel.PeerExchange); None -> NCCL.  The ranks AGREE on the outcome (one
        small all-reduce), so a rank whose setup failed cannot end up in a different collective from its peers."""
        if self._peer_ex is False:
            return None
        if self._peer_ex is None:
            import warnings
            import torch.distributed as dist
            ex, why = None, None
            try:
                if dist.get_backend(pg) != "nccl":
                    raise RuntimeError("not an NCCL group")
                ex = parallel.PeerExchange(parallel.packed_size(K, D), dev, pg)
            except Exception as e:          # symmetric memory unavailable, out of memory, ...
                why = f"{type(e).__name__}: {e}"
            if not parallel.agree(ex is not None, pg):
                if ex is not None:
                    ex.close()
                warnings.warn("b200vq: NVLink peer exchange unavailable on at least one rank"
                              + (f" (here: {why})" if why else "") + "; using NCCL all_reduce for the codebook statistics")
                self._peer_ex = False
                return None
            self._peer_ex = ex
        return self._peer_ex

    # -- extras ------------------------------------------------------------------------------------
    @torch.no_grad()
    def usage_histogram(self) -> Optional[torch.Tensor]:
        """Code usage counts (K,) of the last forward."""
        return None if self._last_stats is None else self._last_stats[:self._num_embeddings]

    @torch.no_grad()
    def global_stats(self):
        """(loss, perplexity) over all data-parallel ranks for the last backward; None otherwise."""
        if self._global_stats is None:
            return None
        lib = _lib.load()
        tail, n_global = self._global_stats
        K, D = self._num_embeddings, self._embedding_dim
        out = torch.empty(2, dtype=torch.float32, device=tail.device)
        tail = tail.contiguous()
        check(lib.vq_finalize_stats(tail.data_ptr(), tail.data_ptr() + 4 * K, n_global, K, D,
                                    float(self._commitment_cost), out.data_ptr(), out.data_ptr() + 4, _stream()))
        return out[0], out[1]

    def encodings_from_indices(self, indices: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Dense (N, K) one-hot for callers that skipped it in forward (vector_quantizer.py:39-40)."""
        lib = _lib.load()
        idx = self.last_indices if indices is None else indices
        idx = idx.to(torch.int32).contiguous()
        out = torch.empty(idx.shape[0], self._num_embeddings, dtype=torch.float32, device=idx.device)
        check(lib.vq_onehot(idx.data_ptr(), idx.shape[0], self._num_embeddings, out.data_ptr(), _stream()))
        return out

    @classmethod
    def from_reference(cls, ref, **kw) -> "VectorQuantizer":
        """Build from a reference `VectorQuantizer` instance (e.g. out of a torch.load'ed pickle)."""
        new = cls(ref._num_embeddings, ref._embedding_dim, ref._commitment_cost,
                  getattr(ref, "_flag_flatten", True), **kw)
        new._embedding.weight.data = ref._embedding.weight.data.clone()
        new._embedding.weight.requires_grad_(ref._embedding.weight.requires_grad)
        new._train_vq = getattr(ref, "_train_vq", True)
        new.train(ref.training)
        return new.to(ref._embedding.weight.device)


def swap_quantizers(model: nn.Module, **kw) -> int:
    """Replace every reference-style quantizer inside `model` (`_vq` of ConvolutionalVQVAE,
    `rir_model._vq` / `speech_model._vq` of EchoedSpeechReconModel) by the B200 one, in place.
    Returns the number of modules swapped."""
    n = 0
    for name, child in list(model.named_children()):
        if isinstance(child, VectorQuantizer):
            continue
        looks_like_vq = (type(child).__name__ == "VectorQuantizer" and hasattr(child, "_embedding")
                         and hasattr(child, "_commitment_cost"))
        if looks_like_vq:
            setattr(model, name, VectorQuantizer.from_reference(child, **kw))
            n += 1
        else:
            n += swap_quantizers(child, **kw)
    return n
