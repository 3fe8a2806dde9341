"""Data-parallel plumbing of the VectorQuantizer hot path (device-agnostic host logic).

The path shards naturally: rows are independent given the replicated codebook, and rows of batch
item b are contiguous in the flattened view (`vector_quantizer.py:32`, N = B*T, item-major), so a
split of the batch dimension is an exact row partition.  Nothing on the data
path is exchanged; per step each rank contributes ONE packed fp32 buffer

        [ dE (K*D) | usage histogram (K) | sum of squared error (1) ]

to a sum all-reduce (NCCL over NVLink on the GPUs; gloo in the CPU tests).  dE is computed by every
rank with the GLOBAL row count in its scale, so the sum is the gradient of the global-batch loss --
the same value DDP's gradient averaging of per-rank mean losses yields for equal shards.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch


def shard_bounds(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) range of batch items owned by `rank` (ragged batches allowed)."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(inputs: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """The (B_r, D, T) slice of a (B, D, T) latent batch for `rank`: an exact row partition."""
    lo, hi = shard_bounds(inputs.shape[0], rank, world)
    return inputs[lo:hi].contiguous()


def packed_size(K: int, D: int) -> int:
    return K * D + K + 1


def new_packed(K: int, D: int, device) -> torch.Tensor:
    return torch.zeros(packed_size(K, D), dtype=torch.float32, device=device)


def packed_views(packed: torch.Tensor, K: int, D: int):
    """(dE (K,D), hist (K,), sse (1,)) views into the packed buffer."""
    return packed[:K * D].view(K, D), packed[K * D:K * D + K], packed[K * D + K:]


def all_reduce_packed(packed: torch.Tensor, group=None) -> torch.Tensor:
    """The single collective of a data-parallel step."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return packed


def global_row_count(n_rows_local: int, group=None, equal_shards: bool = True) -> int:
    """Rows over all ranks.  With equal shards this needs no communication."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return n_rows_local
    world = dist.get_world_size(group)
    if equal_shards or world == 1:
        return n_rows_local * world
    t = torch.tensor([n_rows_local], dtype=torch.int64)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, group=group)
    return int(t.item())


class PeerExchange:
    """Sum all-reduce of the packed step buffer over NVLink peer memory (include/b200vq.h: vq_dp_*).

    Each rank owns two symmetric RECEIVE buffers (torch symmetric memory; alternating between calls).  A call stores
    the local contribution as 16-byte lines {d0, seq, d1, seq} -- data and sequence number in the same store -- into
    the peers and sums what arrives in rank order: no barrier, bit-identical results on every rank.  The sequence
    number lives in device memory and is advanced by the kernel, so calls can be captured in a CUDA graph and replayed
    (nothing on the host has to stay in step with the device).
    """

    def __init__(self, n_floats: int, device, group=None, spin_limit: int = 0):
        import ctypes
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self.lib = _lib.load()
        self.check = _lib.check
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.n = int(n_floats)
        self.device = torch.device(device)
        lines = int(self.lib.vq_dp_recv_lines(self.world, self.n))
        self.bufs, arrays, mcs = [], [], []
        use_mc = os.environ.get("B200VQ_NVLS", "1") != "0"
        for _ in range(2):
            t = symm_mem.empty(lines * 4, dtype=torch.float32, device=self.device)
            t.zero_()
            hdl = symm_mem.rendezvous(t, self.group)
            self.bufs.append(t)
            arrays.append((ctypes.c_void_p * self.world)(*list(hdl.buffer_ptrs)))
            mcs.append(int(getattr(hdl, "multicast_ptr", 0) or 0) if use_mc else 0)
        self.nvls = all(m != 0 for m in mcs)
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)                      # zero-initialised receive buffers are in place everywhere
        ctx = ctypes.c_void_p()
        self.check(self.lib.vq_dp_create(arrays[0], arrays[1], mcs[0] if self.nvls else None, mcs[1] if self.nvls else None,
                                         self.world, self.rank, self.n, spin_limit, ctypes.byref(ctx)))
        self.ctx = ctx

    def allreduce(self, payload: torch.Tensor, out: torch.Tensor, stream_ptr: int) -> torch.Tensor:
        """out = sum over ranks of payload (n_floats each)."""
        assert payload.numel() == self.n and out.numel() == self.n
        self.check(self.lib.vq_dp_allreduce(self.ctx, payload.data_ptr(), out.data_ptr(), stream_ptr))
        return out

    def status(self, stream_ptr: int = 0):
        """(calls completed, error word) -- synchronises the stream; error bit 0 = a bounded wait expired."""
        import ctypes
        calls, err = ctypes.c_uint32(0), ctypes.c_uint32(0)
        self.check(self.lib.vq_dp_status(self.ctx, ctypes.byref(calls), ctypes.byref(err), stream_ptr))
        return calls.value, err.value

    def close(self):
        if getattr(self, "ctx", None) is not None:
            self.lib.vq_dp_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def agree(ok: bool, group=None) -> bool:
    """True only if `ok` holds on EVERY rank (one small all-reduce): ranks must not pick different collectives."""
    import torch.distributed as dist
    t = torch.tensor([1 if ok else 0], dtype=torch.int32)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return bool(int(t.item()) == 1)
