"""torchrun --nproc-per-node N tools/mgpu_allreduce.py : the NVLink exchange (vq_dp_allreduce) against NCCL on N GPUs.
Values (vs dist.all_reduce), bit-identity across ranks, replay from a CUDA graph (device-side sequence numbers), time."""
import os, sys
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
ex = par.PeerExchange(n, dev)
cur = torch.cuda.current_stream()
ok = True
pay = torch.empty(n, device=dev); out = torch.empty(n, device=dev)
for step in range(12):
    torch.manual_seed(100 * step + rank)
    pay.copy_(torch.randn(n, device=dev))
    ref = pay.clone(); dist.all_reduce(ref)
    ex.allreduce(pay, out, cur.cuda_stream)
    torch.cuda.synchronize()
    err = float((out - ref).abs().max() / ref.abs().max())      # rank-order sum vs NCCL's order
    gathered = [torch.empty_like(out) for _ in range(world)]
    dist.all_gather(gathered, out)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    ok = ok and err < 1e-6 and same
    if rank == 0: print(f"step {step}: max rel err vs NCCL {err:.2e}, bit-identical across ranks: {same}")
# the same call replayed from a CUDA graph: sequence numbers and buffer parity advance on the device
g = torch.cuda.CUDAGraph()
side = torch.cuda.Stream(device=dev); side.wait_stream(cur)
with torch.cuda.stream(side):
    with torch.cuda.graph(g, stream=side):
        ex.allreduce(pay, out, torch.cuda.current_stream().cuda_stream)
cur.wait_stream(side)
for rep in range(5):
    pay.copy_(torch.randn(n, device=dev)); ref = pay.clone(); dist.all_reduce(ref)
    g.replay(); torch.cuda.synchronize()
    err = float((out - ref).abs().max() / ref.abs().max())
    ok = ok and err < 1e-6
    if rank == 0: print(f"graph replay {rep}: max rel err vs NCCL {err:.2e}")
for name, fn in (("vq_dp_allreduce", lambda: ex.allreduce(pay, out, cur.cuda_stream)), ("vq_dp_allreduce (graph)", g.replay), ("nccl all_reduce", lambda: dist.all_reduce(out))):
    for _ in range(20): fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(200): fn()
    b.record(); torch.cuda.synchronize()
    if rank == 0: print(f"{name}: {a.elapsed_time(b) / 200 * 1e3:.1f} us per call ({n * 4 / 1024:.0f} KiB, {world} ranks)")
calls, err = ex.status(cur.cuda_stream)
ok = ok and err == 0
if rank == 0: print("ALLREDUCE TEST", "PASSED" if ok else "FAILED", "| nvls:", ex.nvls, "| calls:", calls, "| error word:", err)
ex.close()
dist.destroy_process_group()
