"""Throughput of the nn.Module path (what a user of the drop-in sees) on the bench workload."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200vq
dev = torch.device("cuda:0")
B, D, T, K = 256, 64, 201, 1024
for enc in (True, False):
    vq = b200vq.VectorQuantizer(K, D, 0.25, return_encodings=enc).to(dev)
    vq._embedding.weight.data.normal_()
    zs = [torch.randn(B, D, T, device=dev, requires_grad=True) for _ in range(13)]
    def step(i):
        z = zs[i % 13]
        z.grad = None; vq._embedding.weight.grad = None
        loss, q, perp, e = vq(z)
        (loss + q.sum()).backward()
    for i in range(20): step(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    n = 300
    for i in range(n): step(i)
    b.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"module path, return_encodings={enc}: {a.elapsed_time(b)/n*1e3:.1f} us/step device, {(t1-t0)/n*1e6:.1f} us/step wall -> {B*T*n/(t1-t0)/1e6:.1f} M vectors/s")

# ---- the same step captured in a CUDA graph (static input buffer) ----
vq = b200vq.VectorQuantizer(K, D, 0.25, return_encodings=True).to(dev)
vq._embedding.weight.data.normal_()
static_z = torch.randn(B, D, T, device=dev, requires_grad=True)
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        static_z.grad = None; vq._embedding.weight.grad = None
        loss, q, perp, e = vq(static_z)
        (loss + q.sum()).backward()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
static_z.grad = None; vq._embedding.weight.grad = None
with torch.cuda.graph(g):
    loss, q, perp, e = vq(static_z)
    (loss + q.sum()).backward()
torch.cuda.synchronize()
ref_loss = float(loss); ref_dE = vq._embedding.weight.grad.clone()
new = torch.randn(B, D, T, device=dev)
static_z.data.copy_(new)
g.replay(); torch.cuda.synchronize()
loss2 = float(loss)
# eager check on the same data
vq2 = b200vq.VectorQuantizer(K, D, 0.25).to(dev); vq2._embedding.weight.data.copy_(vq._embedding.weight.data)
zz = new.clone().requires_grad_(True); l3, q3, _, _ = vq2(zz); (l3 + q3.sum()).backward(); torch.cuda.synchronize()
print("graph replay loss", loss2, "eager loss", float(l3), "dE equal:", torch.allclose(vq._embedding.weight.grad, vq2._embedding.weight.grad, rtol=1e-5, atol=1e-9),
      "dz equal:", torch.allclose(static_z.grad, zz.grad))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); a.record()
n = 300
for i in range(n): g.replay()
b.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"module path in a CUDA graph: {a.elapsed_time(b)/n*1e3:.1f} us/step device, {(t1-t0)/n*1e6:.1f} us/step wall -> {B*T*n/(t1-t0)/1e6:.1f} M vectors/s")
