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
