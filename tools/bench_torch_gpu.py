"""Secondary comparison ("GPU reference"): the reference's formulation of the quantizer step --
dense (N, K) distances, argmin, scatter one-hot, one-hot @ E, the two MSE terms, straight-through, perplexity and the
autograd backward (vector_quantizer.py:29-58) -- written with stock PyTorch ops and run on the same B200 in fp32
("highest": no TF32), next to this repo's module on the same inputs.  Not part of bench.py's contract."""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200vq

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
BETA = 0.25


def stock_step(z, E):
    D = E.shape[1]
    flat = z.view(-1, D)
    dist = (flat.pow(2).sum(1, keepdim=True) + E.pow(2).sum(1)) - 2 * torch.matmul(flat, E.t())
    idx = torch.argmin(dist, dim=1).unsqueeze(1)
    enc = torch.zeros(idx.shape[0], E.shape[0], device=z.device)
    enc.scatter_(1, idx, 1)
    q = torch.matmul(enc, E).view(z.shape)
    loss = F.mse_loss(q, z.detach()) + BETA * F.mse_loss(q.detach(), z)
    q_st = z + (q - z).detach()
    p = enc.mean(0)
    perp = torch.exp(-(p * torch.log(p + 1e-10)).sum())
    (loss + q_st.sum()).backward()
    return loss, perp, enc


def timeit(fn, n_warm, n):
    for i in range(n_warm):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e-3


for name, (B, D, T, K), n in [("rir256 (configs[1])", (256, 64, 201, 1024), 50), ("speech32 (configs[0])", (32, 128, 500, 1024), 100),
                              ("N=65536 K=4096 D=128", (128, 128, 512, 4096), 10)]:
    torch.manual_seed(0)
    E = torch.randn(K, D, device=dev, requires_grad=True)
    zs = [torch.randn(B, D, T, device=dev, requires_grad=True) for _ in range(5)]

    def stock(i):
        z = zs[i % 5]
        z.grad = None; E.grad = None
        stock_step(z, E)

    vq = b200vq.VectorQuantizer(K, D, BETA).to(dev)
    vq._embedding.weight.data.copy_(E.detach())

    def mine(i):
        z = zs[i % 5]
        z.grad = None; vq._embedding.weight.grad = None
        loss, q, perp, enc = vq(z)
        (loss + q.sum()).backward()

    t_stock, t_mine = timeit(stock, 5, n), timeit(mine, 5, n)
    N = B * T
    print(f"{name}: stock PyTorch fp32 on B200 {t_stock*1e6:9.1f} us/step = {N/t_stock/1e6:8.1f} M vectors/s | "
          f"b200vq.VectorQuantizer (eager module) {t_mine*1e6:8.1f} us/step = {N/t_mine/1e6:8.1f} M vectors/s | x{t_stock/t_mine:.1f}")
