"""Timeline of the fused forward kernel (debug): builds libb200vq_trace.so with -DVQ_TRACE, runs one
forward on the bench workload and prints per-role clock64 stamps (in us at the nominal 1.9 GHz) for a few CTAs."""
import ctypes, os, subprocess, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from importlib import import_module
B = import_module("acoustic_locating_vq-vae_b200.build")
L = import_module("acoustic_locating_vq-vae_b200._lib")
so = os.path.join(B.CSRC, "libb200vq_trace.so")
srcs = [os.path.join(B.CSRC, f) for f in B.SOURCES + B.HEADERS]
if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in srcs) or "--build-only" in sys.argv:
    subprocess.run([B._nvcc(), *B.NVCC_FLAGS, "-DVQ_TRACE", "-o", so, os.path.join(B.CSRC, "b200vq.cu")], check=True)
if "--build-only" in sys.argv:       # (cross-compile in the authoring container; the .so travels to the GPU box)
    sys.exit(0)
lib = ctypes.CDLL(so)
for name, (res, args) in L.SIGNATURES.items():
    fn = getattr(lib, name); fn.restype = res; fn.argtypes = args
onehot_on = "--no-onehot" not in sys.argv
force_screen = (1 << 10) if "--screen" in sys.argv else 0
Bn, D, T, K = 256, 64, 201, 1024
for a in sys.argv[1:]:
    if a.startswith("--shape="):
        Bn, D, T, K = [int(v) for v in a.split("=")[1].split(",")]
N = Bn * T
dev = torch.device("cuda:0")
torch.manual_seed(0)
E = torch.randn(K, D, device=dev); z = torch.randn(N, D, device=dev)
e2 = torch.empty(K, device=dev); ehi = torch.empty_like(E); elo = torch.empty_like(E)
q = torch.empty_like(z); idx = torch.empty(N, dtype=torch.int32, device=dev)
oh = torch.empty(N, K, device=dev) if onehot_on else None
stats = torch.empty(K + 3, device=dev)
wsb = lib.vq_workspace_bytes(N, K, D, 1); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
trace = torch.zeros(148 * 8 * 64, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream().cuda_stream
lib.vq_debug_set_trace(trace.data_ptr())
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(3):
    trace.zero_()
    assert lib.vq_prepare_codebook(E.data_ptr(), K, D, e2.data_ptr(), ehi.data_ptr(), elo.data_ptr(), st) == 0
    torch.cuda.synchronize()
    ev0.record()
    rc = lib.vq_forward(z.data_ptr(), E.data_ptr(), e2.data_ptr(), ehi.data_ptr(), elo.data_ptr(), N, K, D, 0.25, (1 if onehot_on else 0) | force_screen,
                        q.data_ptr(), idx.data_ptr(), None if oh is None else oh.data_ptr(), stats.data_ptr(), stats.data_ptr() + 4 * K,
                        stats.data_ptr() + 4 * (K + 1), stats.data_ptr() + 4 * (K + 2), ws.data_ptr(), wsb, st)
    ev1.record()
    assert rc == 0, lib.vq_last_error()
    torch.cuda.synchronize()
t = trace.cpu().numpy().reshape(148, 8, 64)
roles = {0: "epi tiles (wait start | acc full | scanned) x12", 1: "mma  (wait z_ready | got | item issued)", 2: "zpipe(tma issue | z landed | converted)", 3: "epi  (a_ready | first acc | last acc | idx out)",
         4: "fill (issue start | issued | [screen: all landed])", 5: "work (cand got | refined | overflow done | item done)", 6: "tail (workers done | zeros seen | ones patched)", 7: "cta  (start | roles done | after cluster sync)"}
slowest = int(np.argmax(t[:, 7, 2] - t[:, 7, 0]))
for cta in (0, 73, slowest, slowest ^ 1):
    t0 = t[cta, 7, 0]
    print(f"--- CTA {cta} (us since role dispatch, 1.9 GHz nominal) ---")
    for r, name in roles.items():
        vals = [(s, (v - t0) / 1900.0) for s, v in enumerate(t[cta, r]) if v != 0 and not (r == 7 and s >= 3) and not (r == 6 and s >= 32)]
        print(f"  {name:50s}", " ".join(f"{s}:{u:.1f}" for s, u in vals[:40]))
cyc = (t[:, 7, 2] - t[:, 7, 0]).astype(np.float64); ns = (t[:, 7, 4] - t[:, 7, 3]).astype(np.float64)
print("effective SM clock during the kernel: %.0f MHz (median over CTAs)" % np.median(cyc / np.maximum(ns, 1) * 1e3))
ld = t[0::2, 6, :]
print("leader MMA thread: cycles waiting on acc_empty: median %.0f ; on full (E TMA): median %.0f ; residency cycles median %.0f" % (
    np.median(ld[:, 32]), np.median(ld[:, 33]), np.median(cyc)))
print("screen kernel: full-rescan rows per CTA: total %d, max %d ; exact pairs per CTA: mean %.0f max %d" % (
    t[:, 6, 34].sum(), t[:, 6, 34].max(), t[:, 6, 35].mean(), t[:, 6, 35].max()))
slow = int(np.argmax(t[:, 7, 2] - t[:, 7, 0]))
print("slowest CTA:", slow, "full-rescan rows:", int(t[slow, 6, 34]), "pairs:", int(t[slow, 6, 35]))
ends = (t[:, 7, 2] - t[:, 7, 0]) / 1900.0
print("kernel residency per CTA (us): min %.1f  median %.1f  max %.1f" % (ends.min(), np.median(ends), ends.max()))
gt0, gt1 = t[:, 7, 3], t[:, 7, 4]
print("globaltimer: first CTA dispatch -> last CTA end %.1f us ; CTA dispatch spread %.1f us ; CUDA-event time of the launch %.1f us" % (
    (gt1.max() - gt0.min()) / 1e3, (gt0.max() - gt0.min()) / 1e3, ev0.elapsed_time(ev1) * 1e3))
if t[:, 7, 5].max() > 0:
    k0, k1 = t[:, 7, 5], t[:, 7, 6]
    print("globaltimer: first instruction of the first CTA -> role dispatch %.1f us (median per CTA %.1f) ; last CTA end -> after TMEM dealloc %.1f us ; first instruction spread over CTAs %.1f us ; whole kernel first instr -> last instr %.1f us" % (
        (gt0.min() - k0.min()) / 1e3, np.median(gt0 - k0) / 1e3, (k1.max() - gt1.max()) / 1e3, (k0.max() - k0.min()) / 1e3, (k1.max() - k0.min()) / 1e3))
