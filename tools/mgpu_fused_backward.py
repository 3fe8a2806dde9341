"""torchrun --nproc-per-node N tools/test_fused_backward_multigpu.py : vq_backward_allreduce (one kernel) against
vq_backward + vq_allreduce_push on N GPUs: dz bit-identical, reduced buffer equal up to the order of the dE atomics and
bit-identical across ranks; then the time of both."""
import os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200vq
from importlib import import_module
par = import_module("acoustic_locating_vq-vae_b200.parallel")
L = import_module("acoustic_locating_vq-vae_b200._lib")
lib = L.load()
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
N, D, K, BETA = 256 * 201, 64, 1024, 0.25
n = K * D + K + 1
ar = par.PushAllReduce(n, dev)
st = torch.cuda.current_stream().cuda_stream
torch.manual_seed(0)
E = torch.randn(K, D, device=dev)
g_loss = torch.full((), 0.7, device=dev)
P = lambda t: t.data_ptr()
ok = True
for step in range(6):
    torch.manual_seed(100 * step + rank)
    z = torch.randn(N, D, device=dev); g = torch.randn(N, D, device=dev)
    idx = torch.randint(0, K, (N,), device=dev, dtype=torch.int32)
    tail = torch.randn(K + 1, device=dev)
    pay = ar.payload()
    # separate kernels
    pay.zero_(); pay[K * D:] = tail
    dz_a = torch.empty_like(z)
    L.check(lib.vq_backward(P(g), P(g_loss), P(z), P(E), P(idx), N, N, N * world, K, D, BETA, L.FLAG_TRAIN_VQ, P(dz_a), P(pay), st))
    out_a = ar.reduce(st).clone()
    # fused kernel
    pay.zero_(); pay[K * D:] = tail
    dz_b = torch.empty_like(z)
    out_b = ar.backward_reduce(P(g), P(g_loss), P(z), P(E), P(idx), N, N * world, K, D, BETA, L.FLAG_TRAIN_VQ, P(dz_b), st).clone()
    torch.cuda.synchronize()
    same_dz = torch.equal(dz_a, dz_b)
    err = float((out_a - out_b).abs().max() / out_a.abs().max())
    tail_ref = tail.clone(); dist.all_reduce(tail_ref)
    tail_err = float((out_b[K * D:] - tail_ref).abs().max())
    gathered = [torch.empty_like(out_b) for _ in range(world)]
    dist.all_gather(gathered, out_b)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    ok = ok and same_dz and err < 1e-5 and same and tail_err < 1e-5
    if rank == 0:
        print(f"step {step}: dz bit-identical {same_dz}; reduced buffer max rel diff {err:.1e}; tail err {tail_err:.1e}; identical across ranks {same}")

z = torch.randn(N, D, device=dev); g = torch.randn(N, D, device=dev)
idx = torch.randint(0, K, (N,), device=dev, dtype=torch.int32); dz = torch.empty_like(z)
pay = ar.payload()


def separate():
    L.check(lib.vq_backward(P(g), P(g_loss), P(z), P(E), P(idx), N, N, N * world, K, D, BETA, L.FLAG_TRAIN_VQ, P(dz), P(pay), st))
    ar.reduce(st)


def fused():
    ar.backward_reduce(P(g), P(g_loss), P(z), P(E), P(idx), N, N * world, K, D, BETA, L.FLAG_TRAIN_VQ, P(dz), st)


for name, fn in (("vq_backward + vq_allreduce_push", separate), ("vq_backward_allreduce (fused)", fused)):
    for _ in range(20): fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(200): fn()
    b.record(); torch.cuda.synchronize()
    if rank == 0: print(f"{name}: {a.elapsed_time(b) / 200 * 1e3:.1f} us per call ({world} ranks)")
if rank == 0: print("FUSED BACKWARD TEST", "PASSED" if ok else "FAILED", "| nvls:", ar.nvls)
dist.destroy_process_group()
