"""World-size-2 `gloo` test (CPU) of the data-parallel host logic: batch sharding is an exact row
partition, and ONE sum all-reduce of the packed [dE | hist | sse] buffer reproduces the
single-process result.  Per-rank arithmetic comes from the CPU oracle; what is
under test is acoustic_locating_vq-vae_b200/parallel.py, the code the CUDA module uses."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, B, D, T, K, ragged, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from importlib import import_module
    par = import_module("acoustic_locating_vq-vae_b200.parallel")
    from oracle import c_oracle
    torch.manual_seed(0)
    E = torch.randn(K, D)
    z = torch.randn(B, D, T)
    g = torch.randn(B, D, T)
    zs, gs = par.shard_batch(z, rank, world), par.shard_batch(g, rank, world)
    rows = zs.numpy().reshape(-1, D)
    n_local = rows.shape[0]
    n_global = par.global_row_count(n_local, equal_shards=not ragged)
    assert n_global == B * T
    idx = c_oracle.argmin(rows, E.numpy())
    fwd = c_oracle.quantize(rows, E.numpy(), idx, 0.25)
    dz, dE = c_oracle.backward(gs.numpy().reshape(-1, D), 1.0, rows, E.numpy(), idx, 0.25, True,
                               n_rows_dz=n_global, n_rows_dE=n_global)
    packed = par.new_packed(K, D, "cpu")
    v_dE, v_hist, v_sse = par.packed_views(packed, K, D)
    v_dE.copy_(torch.from_numpy(dE)); v_hist.copy_(torch.from_numpy(fwd["hist"])); v_sse[0] = fwd["sse"]
    par.all_reduce_packed(packed)
    lo, hi = par.shard_bounds(B, rank, world)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), packed=packed.numpy(), idx=idx, dz=dz, lo=lo, hi=hi)
    dist.destroy_process_group()


@pytest.mark.parametrize("B,ragged", [(8, False), (7, True)])
def test_dp2_packed_allreduce_matches_single_process(tmp_path, B, ragged):
    sys.path.insert(0, ROOT)
    from oracle import c_oracle
    D, T, K, world = 16, 13, 32, 2
    port = 29500 + (os.getpid() % 2000) + (1 if ragged else 0)
    mp.spawn(_worker, args=(world, port, B, D, T, K, ragged, str(tmp_path)), nprocs=world, join=True)
    torch.manual_seed(0)
    E = torch.randn(K, D); z = torch.randn(B, D, T); g = torch.randn(B, D, T)
    rows = z.numpy().reshape(-1, D)
    idx = c_oracle.argmin(rows, E.numpy())
    fwd = c_oracle.quantize(rows, E.numpy(), idx, 0.25)
    dz, dE = c_oracle.backward(g.numpy().reshape(-1, D), 1.0, rows, E.numpy(), idx, 0.25, True)
    r = [np.load(tmp_path / f"rank{i}.npz") for i in range(world)]
    assert np.array_equal(r[0]["packed"], r[1]["packed"])          # every rank holds the same reduced buffer
    packed = r[0]["packed"]
    np.testing.assert_allclose(packed[:K * D].reshape(K, D), dE, rtol=1e-5, atol=1e-9)
    assert np.array_equal(packed[K * D:K * D + K], fwd["hist"])
    assert abs(packed[-1] - fwd["sse"]) <= 1e-5 * fwd["sse"]
    # the shards are an exact partition of the rows: indices and dz concatenate to the full result
    assert int(r[0]["lo"]) == 0 and int(r[0]["hi"]) == int(r[1]["lo"]) and int(r[1]["hi"]) == B
    assert np.array_equal(np.concatenate([r[0]["idx"], r[1]["idx"]]), idx)
    np.testing.assert_allclose(np.concatenate([r[0]["dz"], r[1]["dz"]]), dz, rtol=1e-6, atol=1e-9)
    # global loss / perplexity from the reduced statistics == single-process values
    n = B * T
    m = np.float32(packed[-1] / (n * D))
    assert abs((m + np.float32(0.25) * m) - fwd["loss"]) <= 1e-5 * fwd["loss"]
    p = packed[K * D:K * D + K] / np.float32(n)
    perp = np.exp(-(p * np.log(p + 1e-10)).sum())
    assert abs(perp - fwd["perplexity"]) <= 1e-5 * fwd["perplexity"]
