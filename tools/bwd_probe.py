"""Back-to-back time of vq_backward's three dE strategies (flat atomics / bucket / private), real code distributions.
    python tools/bwd_probe.py [--out gpurun_out/bwd_probe.json]
Inputs rotate over more than L2; idx = argmin of the reference distances (skewed usage, like the bench)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200vq
from importlib import import_module
L = import_module("acoustic_locating_vq-vae_b200._lib")
lib = b200vq.load_library()
dev = torch.device("cuda:0")
HBM = 6534.0
rows = []
SHAPES = ((51456, 64, 1024), (16000, 128, 1024), (32000, 128, 1024), (1 << 20, 64, 512), (1 << 20, 64, 1024), (1 << 20, 128, 1024),
          (1 << 20, 128, 4096), (1 << 19, 256, 2048), (1 << 17, 64, 1024))
only = [a for a in sys.argv if a.startswith("--only=")]
if only:
    SHAPES = tuple(SHAPES[int(i)] for i in only[0][7:].split(","))
REPS = 50
for a in sys.argv:
    if a.startswith("--reps="):
        REPS = int(a[7:])
for (N, D, K) in SHAPES:
    torch.manual_seed(0)
    E = torch.randn(K, D, device=dev)
    nb = max(2, min(14, int(1.3 * 126e6 / (N * D * 4)) + 1))
    zs = [torch.randn(N, D, device=dev) for _ in range(nb)]
    gs = [torch.randn(N, D, device=dev) for _ in range(nb)]
    idxs = []
    for zz in zs:
        parts = []
        for c in range(0, N, 1 << 16):
            zc = zz[c:c + (1 << 16)]
            parts.append(((zc * zc).sum(1, keepdim=True) + (E * E).sum(1) - 2 * zc @ E.t()).argmin(1).int())
        idxs.append(torch.cat(parts))
    dz = torch.empty(N, D, device=dev); dE = torch.zeros(K, D, device=dev)
    gl = torch.ones((), device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def run(flags, reps=REPS):
        def one(i):
            L.check(lib.vq_backward(gs[i % nb].data_ptr(), gl.data_ptr(), zs[i % nb].data_ptr(), E.data_ptr(), idxs[i % nb].data_ptr(),
                                    N, N, N, K, D, 0.25, flags, dz.data_ptr(), dE.data_ptr(), st))
        for i in range(5):
            one(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(reps):
            one(i)
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps * 1e3
    byts = 4 * (3 * N * D + N + 2 * K * D)
    roof = byts / (HBM * 1e9) * 1e6
    r = {"N": N, "D": D, "K": K, "bytes": byts, "roof_us": round(roof, 2), "default_path": lib.vq_backward_path(N, K, D, 0),
         "dz_only_us": round(run(0), 2)}
    for name, fl in (("flat", L.FLAG_BWD_FLAT), ("private", L.FLAG_BWD_PRIVATE)):
        if fl != L.FLAG_BWD_FLAT and lib.vq_backward_path(N, K, D, fl) == 0:
            continue
        if name == "bucket" and N > (1 << 18):
            continue
        t = run(L.FLAG_TRAIN_VQ | fl)      # dE accumulates (no memset in the timed loop)
        r[name + "_us"] = round(t, 2)
        r[name + "_frac"] = round(roof / t, 3)
    rows.append(r)
    print(json.dumps(r), flush=True)
    del zs, gs, idxs
out = [a for a in sys.argv if a.startswith("--out=")]
if out:
    with open(out[0][6:], "w") as f:
        json.dump(rows, f, indent=1)
