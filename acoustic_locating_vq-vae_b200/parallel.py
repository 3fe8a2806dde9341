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


class SymmetricAllReduce:
    """One-shot sum all-reduce of the packed step buffer over NVLink peer memory (vq_allreduce_sum).

    Two symmetric buffers ([payload | flags], torch symmetric memory) alternate between calls; the kernel does a
    flag barrier and sums the peers' payloads in rank order, so every rank gets bit-identical results.  The
    backward kernel accumulates dE straight into `payload()`; `reduce()` returns the reduced packed buffer.
    Falls back to NCCL (`all_reduce_packed`) when symmetric memory cannot be set up.
    """

    def __init__(self, n_floats: int, device, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self.lib = _lib.load()
        self.check = _lib.check
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.n = n_floats
        self.flag_off = (n_floats + 63) // 64 * 64                   # flags start on a 256-byte boundary
        total = self.flag_off + 64
        self.bufs, self.ptr_arrays = [], []
        for _ in range(2):
            t = symm_mem.empty(total, dtype=torch.float32, device=device)
            t.zero_()
            hdl = symm_mem.rendezvous(t, self.group)
            ptrs = list(hdl.buffer_ptrs)
            import ctypes
            arr = (ctypes.c_void_p * self.world)(*ptrs)
            self.bufs.append(t)
            self.ptr_arrays.append(arr)
        torch.cuda.synchronize(device)
        dist.barrier(self.group)                                     # zero-initialised flags are in place everywhere
        self.out = torch.zeros(n_floats, dtype=torch.float32, device=device)
        self.seq = 0

    def payload(self) -> torch.Tensor:
        """The buffer the NEXT reduce() will contribute (write the local partial result here)."""
        return self.bufs[self.seq & 1][:self.n]

    def reduce(self, stream_ptr: int) -> torch.Tensor:
        which = self.seq & 1
        self.seq += 1
        seq_no = (self.seq + 1) // 2                                 # 1, 1, 2, 2, ...: per-buffer sequence number
        self.check(self.lib.vq_allreduce_sum(self.ptr_arrays[which], self.world, self.rank, self.flag_off, self.n,
                                             seq_no, self.out.data_ptr(), stream_ptr))
        return self.out


class PushAllReduce:
    """Low-latency push all-reduce of the packed step buffer over NVLink peer memory (vq_allreduce_push).

    Each rank owns two symmetric RECEIVE buffers (alternating between calls) of `world` slots; a call stores the
    local payload -- data and sequence number in the same 16-byte line -- into slot [rank] of every rank's
    receive buffer and then sums its own slots in rank order as the lines arrive: one NVLink one-way latency,
    no barrier, bit-identical results on every rank.  `payload` is ordinary device memory (the backward kernel
    accumulates dE straight into it).
    """

    def __init__(self, n_floats: int, device, group=None):
        import ctypes
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self.lib = _lib.load()
        self.check = _lib.check
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.n = n_floats
        lines = (n_floats + 1) // 2
        self.bufs, self.ptr_arrays, self.mc_ptrs = [], [], []
        use_mc = os.environ.get("B200VQ_NVLS", "1") != "0"
        for _ in range(2):
            t = symm_mem.empty(self.world * (lines + 2) * 4, dtype=torch.float32, device=device)   # include/b200vq.h: world x (lines + 2) lines
            t.zero_()
            hdl = symm_mem.rendezvous(t, self.group)
            self.bufs.append(t)
            self.ptr_arrays.append((ctypes.c_void_p * self.world)(*list(hdl.buffer_ptrs)))
            mc = int(getattr(hdl, "multicast_ptr", 0) or 0) if use_mc else 0
            self.mc_ptrs.append(mc if mc != 0 else None)
        self.nvls = all(m is not None for m in self.mc_ptrs)
        if not self.nvls:
            self.mc_ptrs = [None, None]
        torch.cuda.synchronize(device)
        dist.barrier(self.group)                      # zero-initialised receive buffers are in place everywhere
        self.payload_buf = torch.zeros(n_floats, dtype=torch.float32, device=device)
        self.out = torch.zeros(n_floats, dtype=torch.float32, device=device)
        self.grid_sync = torch.zeros(2, dtype=torch.int32, device=device)     # grid barrier of the fused backward kernel
        self.seq = 0

    def payload(self) -> torch.Tensor:
        return self.payload_buf

    def _next(self):
        which = self.seq & 1
        self.seq += 1
        return which, (self.seq + 1) // 2             # 1, 1, 2, 2, ...: per-buffer sequence number (never 0)

    def backward_reduce(self, g_q_ptr, g_loss_ptr, z_ptr, E_ptr, idx_ptr, n_rows: int, n_rows_dE: int, K: int, D: int,
                        beta: float, flags: int, dz_ptr: int, stream_ptr: int) -> torch.Tensor:
        """vq_backward + all-reduce in one kernel (vq_backward_allreduce): dE accumulates into payload()[:K*D] (which the
        caller has zeroed, with the histogram / squared error already behind it) and the NVLink exchange overlaps the
        dz pass.  Returns the reduced buffer, like reduce()."""
        which, seq_no = self._next()
        self.check(self.lib.vq_backward_allreduce(g_q_ptr, g_loss_ptr, z_ptr, E_ptr, idx_ptr, n_rows, max(n_rows, 1), n_rows_dE,
                                                  K, D, beta, flags, dz_ptr, self.payload_buf.data_ptr(), self.n,
                                                  self.ptr_arrays[which], self.mc_ptrs[which], self.world, self.rank, seq_no,
                                                  self.grid_sync.data_ptr(), self.out.data_ptr(), stream_ptr))
        return self.out

    def reduce(self, stream_ptr: int) -> torch.Tensor:
        which, seq_no = self._next()
        self.check(self.lib.vq_allreduce_push(self.ptr_arrays[which], self.mc_ptrs[which], self.world, self.rank,
                                              self.payload_buf.data_ptr(), self.n, seq_no, self.out.data_ptr(), stream_ptr))
        return self.out
