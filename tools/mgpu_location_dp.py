"""torchrun --nproc-per-node N tools/mgpu_location_dp.py [--steps K] [--dense]
BASELINE.json configs[4]: the train_location.py pipeline data-parallel on N GPUs, per-GPU batch 16 (train_location.py:25).

What the reference step does (train_location.py:63-94, location_model.py:7-29), restated on synthetic latents of the
shapes the frozen encoders hand over (the conv stacks are stock cuDNN and out of scope; the reference classes cannot
travel to the GPU box):
    RIR-side quantizer (frozen, eval)  z (16, 64, 201)   -> encodings (3216, 1024) one-hot -> reshape (16, 201, 1024)
    speech-side quantizer (frozen)     z (16, 128, 500)  -> results unused (:71)
    LocationModule: fc_1 (201*1024 -> 1024), ReLU, 1024 -> 512 -> 512 -> 64 -> 1;  loss = mse(location, theta / pi); Adam 1e-3
Two variants of the head's first layer:
    default   b200vq.OneHotLinear on the int32 code indices (quantizer with return_encodings=False): the 843 MB fc_1
              weight sees T rows per sample, its gradient is row-sparse (B*T rows) and is all-reduced as (rows, values)
    --dense   the reference formulation: dense one-hot -> nn.Linear(205 824, 1024), dense 843 MB gradient all-reduce
Checks: every rank holds the same parameters after the steps; both variants produce the same loss on step 0.
Prints one JSON line (rank 0) with the step time (CUDA events, max over ranks).
"""
import json, os, sys
import torch, torch.nn as nn, torch.nn.functional as F, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200vq

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
steps = 20
for a in sys.argv[1:]:
    if a.startswith("--steps="):
        steps = int(a[8:])
dense = "--dense" in sys.argv
B, T, K, D_RIR, D_SP, T_SP = 16, 201, 1024, 64, 128, 500

torch.manual_seed(0)                                   # identical parameters on every rank
rir_vq = b200vq.VectorQuantizer(K, D_RIR, 0.25, return_encodings=dense).to(dev)
sp_vq = b200vq.VectorQuantizer(K, D_SP, 0.25, return_encodings=dense).to(dev)
for vq in (rir_vq, sp_vq):
    vq._embedding.weight.data.normal_()
    vq.set_train_vq(False)                             # echoed_speech_model.py:17-18
    vq.eval()
fc1_ref = nn.Linear(T * K, 1024)                       # location_model.py:10 (same init for both variants)
tail = nn.Sequential(nn.ReLU(), nn.Linear(1024, 512), nn.ReLU(), nn.Linear(512, 512), nn.ReLU(), nn.Linear(512, 64), nn.ReLU(), nn.Linear(64, 1)).to(dev)
if dense:
    fc1 = fc1_ref.to(dev)
    params = list(fc1.parameters()) + list(tail.parameters())
    opt = torch.optim.Adam(params, lr=1e-3)
else:
    fc1 = b200vq.OneHotLinear.from_linear(fc1_ref, T, K, sparse_grad=True).to(dev)
    del fc1_ref
    opt = torch.optim.Adam(list(tail.parameters()) + [fc1.bias], lr=1e-3)
    opt_sparse = torch.optim.SparseAdam([fc1.weight_t], lr=1e-3)

g = torch.Generator(device=dev); g.manual_seed(100 + rank)          # every rank its own batch shard


def step(i, timed=False):
    z_rir = torch.randn(B, D_RIR, T, device=dev, generator=g)
    z_sp = torch.randn(B, D_SP, T_SP, device=dev, generator=g)
    theta = torch.rand(B, 1, device=dev, generator=g)
    opt.zero_grad(set_to_none=True)
    if not dense:
        opt_sparse.zero_grad(set_to_none=True)
    _, q, perp, enc = rir_vq(z_rir)                                  # train_location.py:69
    _, q_s, perp_s, enc_s = sp_vq(z_sp)                              # :71 (unused, as in the reference)
    if dense:
        feat = enc.reshape(B, T, K)                                  # :74
        h = fc1(torch.flatten(feat, start_dim=1))
    else:
        h = fc1(rir_vq.last_indices.view(B, T))
    loc = tail(h)
    loss = F.mse_loss(loc, theta)                                    # :77 (theta / pi in the script)
    loss.backward()
    if world > 1:                                                    # data parallel: average the gradients
        dense_grads = [p.grad for p in (list(tail.parameters()) + ([fc1.bias] if not dense else list(fc1.parameters())))]
        flat = torch.cat([t.flatten() for t in dense_grads if t is not None and not t.is_sparse])
        dist.all_reduce(flat)
        flat /= world
        o = 0
        for t in dense_grads:
            t.copy_(flat[o:o + t.numel()].view_as(t)); o += t.numel()
        if not dense:
            # the row-sparse gradient of the 843 MB weight: gather (rows, values) from every rank instead of reducing 843 MB
            gsp = fc1.weight_t.grad.coalesce()
            rows, vals = gsp.indices()[0].contiguous(), gsp.values().contiguous()
            n = torch.tensor([rows.numel()], device=dev)
            ns = [torch.zeros_like(n) for _ in range(world)]
            dist.all_gather(ns, n)
            m = int(max(int(x) for x in ns))
            rows_p = torch.full((m,), -1, dtype=rows.dtype, device=dev); rows_p[:rows.numel()] = rows
            vals_p = torch.zeros(m, vals.shape[1], device=dev); vals_p[:vals.shape[0]] = vals
            all_r = [torch.empty_like(rows_p) for _ in range(world)]; all_v = [torch.empty_like(vals_p) for _ in range(world)]
            dist.all_gather(all_r, rows_p); dist.all_gather(all_v, vals_p)
            R = torch.cat(all_r); V = torch.cat(all_v) / world
            keep = R >= 0
            fc1.weight_t.grad = torch.sparse_coo_tensor(R[keep][None, :], V[keep], fc1.weight_t.shape)
    opt.step()
    if not dense:
        opt_sparse.step()
    return loss.detach()


for i in range(3):
    l0 = step(i)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(steps):
    loss = step(i)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / steps
chk = torch.stack([p.detach().double().sum() for p in tail.parameters()]).sum() + (fc1.bias.detach().double().sum() if fc1.bias is not None else 0)
same = True
if world > 1:
    t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t[0])
    cs = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(cs, chk)
    same = all(abs(float(c - cs[0])) <= 1e-9 * max(1.0, abs(float(cs[0]))) for c in cs)
if rank == 0:
    print(json.dumps({"config": "configs[4]: train_location.py pipeline, per-GPU batch 16, data parallel", "n_gpus": world,
                      "first_layer": "dense one-hot -> nn.Linear (reference formulation)" if dense else "OneHotLinear on code indices, row-sparse gradient exchange",
                      "ms_per_step": round(ms, 3), "samples_per_s": round(B * world / (ms * 1e-3)), "loss": float(loss),
                      "parameters_identical_on_all_ranks": bool(same), "steps": steps}))
if world > 1:
    dist.destroy_process_group()
