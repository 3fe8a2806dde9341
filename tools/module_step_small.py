"""A handful of eager nn.Module steps on the bench workload (run under `ncu --metrics gpu__time_duration.sum` for the launch
list of one module step: which kernels besides the library's run, and for how long)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200vq
dev = torch.device("cuda:0")
B, D, T, K = 256, 64, 201, 1024
vq = b200vq.VectorQuantizer(K, D, 0.25).to(dev)
vq._embedding.weight.data.normal_()
z = torch.randn(B, D, T, device=dev, requires_grad=True)
gq = torch.ones(B, D, T, device=dev)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    z.grad = None; vq._embedding.weight.grad = None
    loss, q, perp, enc = vq(z)
    torch.autograd.backward([loss, q], [None, gq])
torch.cuda.synchronize()
print("ok", float(loss))
