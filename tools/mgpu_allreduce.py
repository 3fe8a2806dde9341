"""torchrun --nproc-per-node N tools/test_allreduce_multigpu.py : vq_allreduce_sum vs NCCL on N GPUs."""
import os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200vq
from importlib import import_module
par = import_module("acoustic_locating_vq-vae_b200.parallel")
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = 1024 * 64 + 1024 + 1
ar = (par.SymmetricAllReduce if "--pull" in sys.argv else par.PushAllReduce)(n, dev)
st = torch.cuda.current_stream().cuda_stream
ok = True
for step in range(12):
    torch.manual_seed(100 * step + rank)
    x = torch.randn(n, device=dev)
    ar.payload().copy_(x)
    ref = x.clone(); dist.all_reduce(ref)
    out = ar.reduce(st)
    torch.cuda.synchronize()
    # rank-order sum is deterministic and identical on all ranks; NCCL may use another order
    err = float((out - ref).abs().max() / ref.abs().max())
    gathered = [torch.empty_like(out) for _ in range(world)]
    dist.all_gather(gathered, out)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    ok = ok and err < 1e-6 and same
    if rank == 0: print(f"step {step}: max rel err vs NCCL {err:.2e}, bit-identical across ranks: {same}")
# timing
for name, fn in (("vq_allreduce_sum", lambda: ar.reduce(st)), ("nccl all_reduce", lambda: dist.all_reduce(ar.out))):
    for _ in range(20): fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(200): fn()
    b.record(); torch.cuda.synchronize()
    if rank == 0: print(f"{name}: {a.elapsed_time(b) / 200 * 1e3:.1f} us per call ({n * 4 / 1024:.0f} KiB, {world} ranks)")
if rank == 0: print("ALLREDUCE TEST", "PASSED" if ok else "FAILED", "| nvls:", getattr(ar, "nvls", None))
dist.destroy_process_group()
