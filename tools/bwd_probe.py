"""Times vq_backward variants on the bench workload: with / without the dE scatter-add, with / without g_q."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200vq
from importlib import import_module
L = import_module("acoustic_locating_vq-vae_b200._lib")
lib = b200vq.load_library()
dev = torch.device("cuda:0")
for (B, D, T, K) in ((256, 64, 201, 1024), (1024, 64, 1024, 1024), (32, 128, 500, 1024)):
    N = B * T
    torch.manual_seed(0)
    E = torch.randn(K, D, device=dev)
    nb = max(3, int(1.3 * 126e6 / (N * D * 4)) + 1)
    zs = [torch.randn(N, D, device=dev) for _ in range(nb)]
    gs = [torch.randn(N, D, device=dev) for _ in range(nb)]
    idx = torch.randint(0, K, (N,), dtype=torch.int32, device=dev)
    idx_sorted = torch.sort(idx).values.contiguous()
    dz = torch.empty(N, D, device=dev); dE = torch.zeros(K, D, device=dev)
    gl = torch.ones((), device=dev)
    st = torch.cuda.current_stream().cuda_stream
    def run(flags, g, ix, reps=100):
        for i in range(5):
            L.check(lib.vq_backward(None if g is None else gs[i % nb].data_ptr(), gl.data_ptr(), zs[i % nb].data_ptr(), E.data_ptr(), ix.data_ptr(), N, N, N, K, D, 0.25, flags, dz.data_ptr(), dE.data_ptr(), st))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(reps):
            L.check(lib.vq_backward(None if g is None else gs[i % nb].data_ptr(), gl.data_ptr(), zs[i % nb].data_ptr(), E.data_ptr(), ix.data_ptr(), N, N, N, K, D, 0.25, flags, dz.data_ptr(), dE.data_ptr(), st))
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps * 1e3
    print(f"N={N} D={D} K={K}: dz only {run(0, True, idx):.1f} us | dz+dE {run(L.FLAG_TRAIN_VQ, True, idx):.1f} us | dz+dE+memset {run(L.FLAG_TRAIN_VQ | L.FLAG_ZERO_DE, True, idx):.1f} us | "
          f"dz+dE sorted idx {run(L.FLAG_TRAIN_VQ, True, idx_sorted):.1f} us | no g_q, dE {run(L.FLAG_TRAIN_VQ, None, idx):.1f} us | "
          f"bytes {4*(3*N*D+N)/1e6:.1f} MB -> {4*(3*N*D+N)/6.5e12*1e6:.1f} us at 6.5 TB/s")
