"""Where the eager nn.Module path spends its host time (cProfile over 300 steps of the RIR-256 workload)."""
import cProfile, os, pstats, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200vq
dev = torch.device("cuda:0")
B, D, T, K = 256, 64, 201, 1024
vq = b200vq.VectorQuantizer(K, D, 0.25).to(dev)
vq._embedding.weight.data.normal_()
zs = [torch.randn(B, D, T, device=dev, requires_grad=True) for _ in range(13)]
g = torch.ones(B, D, T, device=dev)


def step(i):
    z = zs[i % 13]
    z.grad = None; vq._embedding.weight.grad = None
    loss, q, perp, e = vq(z)
    torch.autograd.backward([loss, q], [torch.ones_like(loss), g])


for i in range(30):
    step(i)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for i in range(300):
    step(i)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host time per step {1e6*(t1-t0)/300:.1f} us ; with final sync {1e6*(t2-t0)/300:.1f} us")
pr = cProfile.Profile()
pr.enable()
for i in range(300):
    step(i)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
