"""torchrun --nproc-per-node N tools/mgpu_module_dp.py : the nn.Module under data parallelism on N GPUs.
Every rank quantizes its batch shard; dE / global loss / perplexity must equal the single-process oracle on the
full batch, and dE must be bit-identical on every rank."""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200vq
from importlib import import_module
par = import_module("acoustic_locating_vq-vae_b200.parallel")
from oracle import c_oracle
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
B, D, T, K = 8 * world, 64, 201, 1024
ok = True
for exact in (True, False):
    torch.manual_seed(0)
    E = torch.randn(K, D); z = torch.randn(B, D, T); g = torch.randn(B, D, T)
    vq = b200vq.VectorQuantizer(K, D, 0.25, exact=exact, data_parallel=True).to(dev)
    vq._embedding.weight.data.copy_(E)
    for step in range(3):
        vq._embedding.weight.grad = None
        zs = par.shard_batch(z, rank, world).to(dev).requires_grad_(True)
        gs = par.shard_batch(g, rank, world).to(dev)
        loss, q, perp, enc = vq(zs)
        (loss + (gs * q).sum()).backward()
        torch.cuda.synchronize()
    rows = z.numpy().reshape(-1, D)
    idx = c_oracle.argmin(rows, E.numpy())
    fwd = c_oracle.quantize(rows, E.numpy(), idx, 0.25)
    N = rows.shape[0]
    dz_ref, dE_ref = c_oracle.backward(g.numpy().reshape(-1, D), 1.0, rows, E.numpy(), idx, 0.25, True, n_rows_dz=N // world, n_rows_dE=N)
    dE = vq._embedding.weight.grad
    gl, gp = vq.global_stats()
    e1 = float(np.abs(dE.cpu().numpy() - dE_ref).max() / np.abs(dE_ref).max())
    lo, hi = par.shard_bounds(B, rank, world)
    e2 = float(np.abs(zs.grad.cpu().numpy().reshape(-1, D) - dz_ref[lo * T:hi * T]).max())
    e3 = abs(float(gl) - fwd["loss"]) / fwd["loss"]; e4 = abs(float(gp) - fwd["perplexity"]) / fwd["perplexity"]
    gathered = [torch.empty_like(dE) for _ in range(world)]
    dist.all_gather(gathered, dE)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    used_push = vq._peer_ex not in (None, False)
    good = e1 < 1e-5 and e2 < 1e-6 and e3 < 1e-5 and e4 < 1e-5 and same
    ok = ok and good
    if rank == 0:
        print(f"exact={exact}: dE rel err {e1:.1e}, dz abs err {e2:.1e}, global loss rel err {e3:.1e}, perplexity rel err {e4:.1e}, "
              f"dE bit-identical across {world} ranks: {same}, NVLink exchange used: {used_push} -> {'ok' if good else 'FAIL'}")
if rank == 0: print("MODULE DP TEST", "PASSED" if ok else "FAILED")
dist.destroy_process_group()
