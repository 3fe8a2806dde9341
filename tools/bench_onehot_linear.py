"""fc_1(one_hot) (location_model.py:10,21; train_location.py:74-75) -- dense reference path vs the index gather."""
import os, sys
import torch, torch.nn as nn, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200vq
dev = torch.device("cuda:0")
B, T, K, O = 16, 201, 1024, 1024
torch.manual_seed(0)
lin = nn.Linear(T * K, O).to(dev)
m = b200vq.OneHotLinear.from_linear(lin, T, K)
idx = torch.randint(0, K, (B, T), device=dev, dtype=torch.int32)
def timed(fn, reps=50):
    for _ in range(5): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
def dense_fwd():
    oh = F.one_hot(idx.long(), K).float()
    return lin(torch.flatten(oh, start_dim=1))
def dense_fwd_bwd():
    y = dense_fwd(); lin.zero_grad(set_to_none=True); y.sum().backward()
def gather_fwd():
    return m(idx)
def gather_fwd_bwd():
    y = m(idx); m.zero_grad(set_to_none=True); y.sum().backward()
print(f"B={B} T={T} K={K} O={O}: weight {T*K*O*4/1e6:.0f} MB")
print(f"dense   fc_1(one_hot): fwd {timed(dense_fwd):8.1f} us   fwd+bwd {timed(dense_fwd_bwd):8.1f} us")
print(f"gather  OneHotLinear : fwd {timed(gather_fwd):8.1f} us   fwd+bwd {timed(gather_fwd_bwd):8.1f} us (row-sparse weight gradient)")
print("max |diff| fwd:", float((dense_fwd() - gather_fwd()).abs().max()))
