#!/usr/bin/env python
"""bench.py -- quantized vectors/s for the VectorQuantizer hot path (forward + backward).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

One "step" = one pass of the hot path (vq_prepare_codebook + vq_forward + vq_backward, i.e.
vector_quantizer.py:29-58 and its autograd) over one batch of synthetic latents of the shape
`_pre_vq_conv` hands to the quantizer.  Default workload = BASELINE.json configs[1]:
RIR VQ-VAE quantizer from train_rir.py defaults at batch 256 -> z (256, 64, 201), N = 51 456 rows,
K = 1024, D = 64, beta = 0.25, dense one-hot `encodings` emitted (the reference always returns it).

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline      dominant kernel: algorithmic bytes|flops per launch / CUDA-event time per launch
  cpu_baseline  the oracle port (same aten ops as the reference module) timed on this box's host cores
  e2e           same metric through the host-buffer C ABI (pinned host z in, loss/perplexity/indices out)
  kernels       per-kernel share of the step (CUDA events on the launching stream)
Under torchrun (N > 1) every rank runs the same per-rank workload on its own rows (weak scaling) and
all-reduces the packed [dE | hist | sse] buffer once per step (own one-shot NVLink kernel; --nccl for NCCL).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (B, D, T, K, description)
    "rir256": (256, 64, 201, 1024, "configs[1]: RIR VQ-VAE quantizer (train_rir.py defaults), batch 256"),
    "speech32": (32, 128, 500, 1024, "configs[0]: speech VQ-VAE quantizer (train_speech.py defaults), batch 32"),
    "echoed64": (64, 128, 500, 1024, "configs[2] speech side: echoed-speech step, batch 64"),
    "sweep_k1024_d64": (1024, 64, 1024, 1024, "configs[3]: N=1M rows, K=1024, D=64"),
    "sweep_k4096_d128": (1024, 128, 1024, 4096, "configs[3]: N=1M rows, K=4096, D=128"),
    "sweep_k8192_d128": (1024, 128, 1024, 8192, "configs[3]: N=1M rows, K=8192, D=128"),
    "sweep_k512_d64": (1024, 64, 1024, 512, "configs[3]: N=1M rows, K=512, D=64"),
    "sweep_k2048_d256": (1024, 256, 1024, 2048, "configs[3]: N=1M rows, K=2048, D=256"),
    "sweep_k8192_d256": (512, 256, 1024, 8192, "configs[3]: N=512k rows, K=8192, D=256"),
}
BETA = 0.25
L2_BYTES = 126 * 1024 * 1024
KERNEL_NAMES = ["prepare_codebook", "argmin_tc", "argmin_exact", "rows", "backward", "finalize", "onehot", "allreduce"]


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d["bf16_tflops"]),
                    bf16_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_sustained=1400.0, source="fallback")


def load_traffic():
    """dram bytes per launch of each kernel from the committed ncu --set full capture (or None)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {}


class ClockSampler:
    """Samples SM clock and throttle reasons with NVML while the timed regions run."""

    def __init__(self, index: int, period_s: float = 0.01):
        self.samples = []
        self.reasons = 0
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._h = None
        self.period = period_s

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)
                util = nv.nvmlDeviceGetUtilizationRates(self._h).gpu
                self.samples.append((mhz, util))
                self.reasons |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
            except Exception:
                try:
                    self.reasons |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                except Exception:
                    pass
            time.sleep(self.period)

    def start(self):
        if self._h is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        busy = [m for m, u in self.samples if u > 0] or [m for m, _ in self.samples]
        names = []
        bits = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}
        for b, n in bits.items():
            if self.reasons & b:
                names.append(n)
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": self.max_mhz, "reasons": names,
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation of the path.  The reference is a Python module
# that cannot travel to the GPU box, so this times the oracle port (oracle/vq_oracle.py: the same aten
# calls as vector_quantizer.py:29-58, bit-identical to it on CPU) with all host threads.
# ------------------------------------------------------------------------------------------------------
def cpu_reference_run(B, D, T, K, steps, warmup, budget_s=150.0):
    import torch
    from oracle import vq_oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    E = torch.randn(K, D)
    z = torch.randn(B, D, T)
    t0 = time.perf_counter()
    vq_oracle.forward_backward_dense(z, E, BETA)
    t_probe = time.perf_counter() - t0
    b_eff = B
    total = (steps + warmup) * t_probe
    if total > budget_s:
        b_eff = max(1, int(B * budget_s / total))
    zs = z[:b_eff].contiguous()
    for _ in range(warmup):
        vq_oracle.forward_backward_dense(zs, E, BETA)
    t0 = time.perf_counter()
    for _ in range(steps):
        vq_oracle.forward_backward_dense(zs, E, BETA)
    dt = time.perf_counter() - t0
    rows = b_eff * T
    return dict(value=rows * steps / dt, ms_per_step=1e3 * dt / steps, cores=cores, rows_per_step=rows,
                sample=f"{b_eff} of {B} batch items per step ({rows} rows), {steps} timed steps, "
                       f"torch {torch.__version__} CPU, {cores} threads")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="rir256", choices=sorted(WORKLOADS))
    ap.add_argument("--no-onehot", action="store_true", help="indices-only mode (encodings not materialised)")
    ap.add_argument("--exact", action="store_true", help="CUDA-core exact path instead of tcgen05")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-fuse", action="store_true", help="argmin and row epilogue as two kernels (VQ_FLAG_NO_FUSE)")
    ap.add_argument("--no-screen", action="store_true", help="3xTF32 tensor kernels instead of screen + exact refine (VQ_FLAG_NO_SCREEN)")
    ap.add_argument("--screen", action="store_true", help="force screen + exact refine (VQ_FLAG_SCREEN)")
    ap.add_argument("--split-backward", action="store_true", help="N > 1: dE-only backward, all-reduce on a side stream while the dz pass runs (measured slower than the serial default: cross-stream events cost more than the overlap gains)")
    ap.add_argument("--fused-allreduce", action="store_true", help="N > 1: vq_backward_allreduce (one kernel, exchange overlaps the dz pass; measured slower) instead of vq_backward + vq_allreduce_push")
    ap.add_argument("--nccl", action="store_true", help="N > 1: use NCCL for the per-step all-reduce instead of vq_allreduce_sum")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    B, D, T, K, desc = WORKLOADS[args.workload]
    N = B * T
    emit_onehot = not args.no_onehot and N * K * 4 <= (8 << 30)

    # -------------------------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        r = cpu_reference_run(B, D, T, K, args.steps, args.warmup)
        line = {
            "impl": "reference", "metric": "quantized vectors/sec (VQ fwd+bwd)", "value": r["value"],
            "unit": "vectors/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "B": B, "D": D, "T": T, "K": K, "rows_per_step": r["rows_per_step"],
                       "beta": BETA, "encodings": "dense one-hot (as the reference)"},
            "cpu_baseline": {"value": r["value"], "unit": "vectors/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "vectors/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    # -------------------------------------------------------------------------------- B200 arm
    import torch
    import b200vq
    from importlib import import_module
    L = import_module("acoustic_locating_vq-vae_b200._lib")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a B200 (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    b200vq.build_extension()
    lib = b200vq.load_library()
    L.check(lib.vq_device_check())
    peaks = load_peaks()

    torch.manual_seed(1000 + rank)
    g = torch.Generator(device=dev)
    g.manual_seed(1000 + rank)
    torch.manual_seed(0)
    E = torch.randn(K, D).to(dev)                                     # same codebook on every rank
    nbuf = max(3, int(1.25 * L2_BYTES / (N * D * 4)) + 1)             # rotating inputs: set larger than L2
    zs = [torch.randn(N, D, device=dev, generator=g) for _ in range(nbuf)]
    gs = [torch.randn(N, D, device=dev, generator=g) for _ in range(nbuf)]
    e2 = torch.empty(K, device=dev); ehi = torch.empty(K, D, device=dev); elo = torch.empty(K, D, device=dev)
    q = torch.empty(N, D, device=dev); idx = torch.empty(N, dtype=torch.int32, device=dev)
    onehot = torch.empty(N, K, device=dev) if emit_onehot else None
    # [dE | hist | sse] -> one all-reduce per step.  With N > 1 the buffer lives in symmetric (peer-mapped) memory and
    # is reduced by our own one-shot NVLink kernel (vq_allreduce_sum); NCCL is the fallback.
    n_packed = K * D + K + 1
    sym = None
    collective = "none"
    if world > 1 and not args.nccl:
        try:
            par = import_module("acoustic_locating_vq-vae_b200.parallel")
            sym = par.PushAllReduce(n_packed, dev)
            collective = "low-latency push all-reduce over NVLink peer memory (vq_allreduce_push, " + ("NVLS multimem.st" if sym.nvls else "P2P stores") + ")"
        except Exception as e:      # symmetric memory unavailable: keep going with NCCL
            sym = None
            collective = f"NCCL all_reduce (symmetric memory unavailable: {type(e).__name__})"
    elif world > 1:
        collective = "NCCL all_reduce"
    fused_ar = sym is not None and D % 4 == 0 and args.fused_allreduce
    split_bwd = sym is not None and not fused_ar and args.split_backward
    main_stream = torch.cuda.current_stream()
    side_stream = torch.cuda.Stream(device=dev) if split_bwd else None
    ev_dE, ev_ar = torch.cuda.Event(), torch.cuda.Event()
    if split_bwd:
        collective = ("backward split in two (dE, then dz) so that the push all-reduce of [dE|hist|sse] over NVLink peer memory "
                      "(vq_allreduce_push, " + ("NVLS multimem.st" if sym.nvls else "P2P stores") + ") runs on a side stream while dz is computed")
    if fused_ar:
        collective = "backward + two-step push all-reduce fused in one kernel (vq_backward_allreduce, " + ("NVLS multimem.st" if sym.nvls else "P2P stores") + ")"
    packs = [sym.payload()] if sym is not None else [torch.zeros(n_packed, device=dev)]
    views = [(pk[:K * D], pk[K * D:K * D + K], pk[K * D + K:]) for pk in packs]
    scal = torch.empty(2, device=dev)                                 # loss, perplexity
    dz = torch.empty(N, D, device=dev)
    g_loss = torch.ones((), device=dev)
    fwd_flags = ((L.FLAG_ONEHOT if emit_onehot else 0) | (L.FLAG_EXACT if args.exact else 0) |
                 (L.FLAG_NO_FUSE if args.no_fuse else 0) | (L.FLAG_NO_SCREEN if args.no_screen else 0) | (L.FLAG_SCREEN if args.screen else 0) | L.FLAG_STATE_READY)
    bwd_flags = L.FLAG_TRAIN_VQ
    # the bucket backward writes every dE element exactly once (VQ_FLAG_ZERO_DE = plain stores); the other paths
    # accumulate with atomics into a buffer that the prepare launch zeroes
    zero_in_prepare = lib.vq_backward_path(N, K, D, 0) != 1
    if not zero_in_prepare:
        bwd_flags |= L.FLAG_ZERO_DE
    wsb = lib.vq_workspace_bytes(N, K, D, fwd_flags)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    n_dE = N * world
    P = lambda t: None if t is None else t.data_ptr()

    def step(i):
        z = zs[i % nbuf]
        dE, hist, sse = views[0]
        # one launch: codebook norms + tf32 split + reset of hist / completion counter / dE accumulator
        L.check(lib.vq_prepare_step(P(E), K, D, P(e2), P(ehi), P(elo), P(hist), P(ws), wsb, P(dE) if zero_in_prepare else None, st))
        L.check(lib.vq_forward(P(z), P(E), P(e2), P(ehi), P(elo), N, K, D, BETA, fwd_flags, P(q), P(idx), P(onehot),
                               P(hist), P(sse), scal.data_ptr(), scal.data_ptr() + 4, P(ws), wsb, st))
        if split_bwd:
            # data parallel: codebook gradient first, then its all-reduce on a side stream WHILE the dz pass runs
            L.check(lib.vq_backward(None, P(g_loss), P(z), P(E), P(idx), N, N, n_dE, K, D, BETA, bwd_flags | L.FLAG_NO_DZ,
                                    None, P(dE), st))
            ev_dE.record(main_stream)
            side_stream.wait_event(ev_dE)
            sym.reduce(side_stream.cuda_stream)          # reduced [dE | hist | sse] lands in sym.out
            ev_ar.record(side_stream)
            L.check(lib.vq_backward(P(gs[i % nbuf]), P(g_loss), P(z), P(E), P(idx), N, N, n_dE, K, D, BETA, 0, P(dz), None, st))
            main_stream.wait_event(ev_ar)                # the step ends when both are done
            return
        if fused_ar:
            # backward + all-reduce in ONE kernel: the NVLink exchange overlaps the dz pass; result in sym.out
            sym.backward_reduce(P(gs[i % nbuf]), P(g_loss), P(z), P(E), P(idx), N, n_dE, K, D, BETA, bwd_flags, P(dz), st)
            return
        L.check(lib.vq_backward(P(gs[i % nbuf]), P(g_loss), P(z), P(E), P(idx), N, N, n_dE, K, D, BETA, bwd_flags,
                                P(dz), P(dE), st))
        if sym is not None:
            sym.reduce(st)                       # reduced [dE | hist | sse] lands in sym.out
        elif world > 1:
            dist.all_reduce(packs[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    # ---- device-resident throughput ------------------------------------------------------------------
    for i in range(args.warmup):
        step(i)
    barrier()
    sampler.start()
    launches0 = lib.vq_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    ev1.record()
    barrier()
    launches = lib.vq_launch_count() - launches0
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    value = N * world * args.steps / (ms * 1e-3)
    loss_val, perp_val = [float(x) for x in scal.tolist()]

    # ---- per-kernel timing (CUDA events on the launching stream) for the roofline ---------------------
    lib.vq_profile_enable(1)
    for i in range(args.steps):
        step(i)
    barrier()
    kern = {}
    for kid, name in enumerate(KERNEL_NAMES):
        tot, cnt = ctypes.c_double(0), ctypes.c_int64(0)
        L.check(lib.vq_profile_read(kid, ctypes.byref(tot), ctypes.byref(cnt)))
        if cnt.value:
            kern[name] = {"launches": cnt.value, "avg_us": 1e3 * tot.value / cnt.value}
    lib.vq_profile_enable(0)
    tot_us = sum(k["avg_us"] * k["launches"] for k in kern.values()) / max(args.steps, 1)
    for k in kern.values():
        k["share"] = round(k["avg_us"] * k["launches"] / args.steps / tot_us, 4)
        k["avg_us"] = round(k["avg_us"], 3)
    dom = max(kern, key=lambda n: kern[n]["avg_us"] * kern[n]["launches"])
    traffic = load_traffic()
    fused = "rows" not in kern and "argmin_tc" in kern          # row epilogue ran inside the tensor kernel
    rows_bytes = 4.0 * (2 * N * D + N + K * D + K + (N * K if emit_onehot else 0))
    # algorithmic work per launch: SURVEY.md section 8(d) per-row figures x rows per launch
    alg = {
        "argmin_tc": {"flops": 2.0 * N * K * D, "bytes": rows_bytes if fused else 4.0 * (N * D + N + 2 * K * D)},
        "argmin_exact": {"flops": 2.0 * N * K * D, "bytes": 4.0 * (N * D + N + K * D)},
        "rows": {"flops": 0.0, "bytes": rows_bytes},
        "backward": {"flops": 0.0, "bytes": 4.0 * (3 * N * D + N + 2 * K * D)},
        "prepare_codebook": {"flops": 0.0, "bytes": 4.0 * (3 * K * D + K)},
    }
    work = alg.get(dom, {"flops": 0.0, "bytes": 0.0})
    dur_s = kern[dom]["avg_us"] * 1e-6
    tf32_peak = peaks["bf16_tflops"] / 2.0                        # TF32 pipe = half the measured dense bf16 rate
    t_tensor = work["flops"] / (tf32_peak * 1e12)
    t_hbm = work["bytes"] / (peaks["hbm_gbs"] * 1e9)
    if t_tensor >= t_hbm:                                          # whichever roofline binds this launch
        bound, peak, achieved, unit = "tensor", tf32_peak, work["flops"] / dur_s / 1e12, "TFLOP/s"
        per_launch = work["flops"]
    else:
        bound, peak, achieved, unit = "hbm", peaks["hbm_gbs"], work["bytes"] / dur_s / 1e9, "GB/s"
        per_launch = work["bytes"]
    screen_used = (not args.exact and not args.no_screen and not args.no_fuse and K % 256 == 0 and D in (32, 64, 96, 128, 192, 256)
                   and os.environ.get("B200VQ_SCREEN", "1")[:1] != "0") or (args.screen and K % 256 == 0)
    fused_name = ("fused forward (vq_screen_kernel: TF32 screen + exact refine + row epilogue)" if screen_used
                  else "fused forward (argmin_tc2: 3xTF32 + row epilogue)")
    roofline = {"kernel": (fused_name if fused and dom == "argmin_tc" else dom), "bound": bound,
                "achieved": round(achieved, 2), "peak": round(peak, 1), "unit": unit, "frac": round(achieved / peak, 4),
                "traffic": traffic.get(args.workload, {}).get(dom),
                "peak_source": peaks["source"] + (" bf16/2 (tf32 pipe)" if bound == "tensor" else " copy bandwidth"),
                "per_launch": per_launch, "avg_us": kern[dom]["avg_us"],
                "other_roof": {"tensor_tflops": round(work["flops"] / dur_s / 1e12, 2), "hbm_gbs": round(work["bytes"] / dur_s / 1e9, 1)}}

    # ---- end to end through the host-buffer C ABI: pinned host z in, loss/perplexity/indices out ------
    e2e = None
    if not args.skip_e2e:
        ctx = ctypes.c_void_p()
        L.check(lib.vq_host_ctx_create(N, K, D, ctypes.byref(ctx)))
        E_host = E.cpu().contiguous()
        L.check(lib.vq_host_set_codebook(ctx, E_host.data_ptr()))
        nhost = 4
        z_host = [torch.randn(N, D).pin_memory() for _ in range(nhost)]
        res = [dict(loss=torch.zeros(1).pin_memory(), perp=torch.zeros(1).pin_memory(),
                    idx=torch.zeros(N, dtype=torch.int32).pin_memory()) for _ in range(2)]
        lane_t = []
        if world > 1:
            for lane in range(2):
                s_, de_, hi_ = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
                L.check(lib.vq_host_lane_buffers(ctx, lane, ctypes.byref(s_), ctypes.byref(de_), ctypes.byref(hi_)))

                class _Arr:
                    pass
                a = _Arr()
                a.__cuda_array_interface__ = {"shape": (K * D,), "typestr": "<f4", "data": (de_.value, False), "version": 2}
                lane_t.append((torch.cuda.ExternalStream(s_.value), torch.as_tensor(a, device=dev)))

        def host_step(i):
            lane = i & 1
            L.check(lib.vq_host_wait(ctx, lane))          # results of step i-2 are now in res[lane]
            r = res[lane]
            L.check(lib.vq_host_step_async(ctx, lane, z_host[i % nhost].data_ptr(), None, N, n_dE, BETA,
                                           L.FLAG_TRAIN_VQ | (L.FLAG_EXACT if args.exact else 0),
                                           r["loss"].data_ptr(), r["perp"].data_ptr(), r["idx"].data_ptr(), None, None, None))
            if world > 1:
                s_, t_ = lane_t[lane]
                with torch.cuda.stream(s_):
                    dist.all_reduce(t_)

        e_steps = args.steps
        for i in range(args.warmup):
            host_step(i)
        L.check(lib.vq_host_wait(ctx, 0)); L.check(lib.vq_host_wait(ctx, 1))
        barrier()
        L.check(lib.vq_host_timer_start(ctx))
        for i in range(e_steps):
            host_step(i)
        e_ms = ctypes.c_float(0)
        L.check(lib.vq_host_timer_stop_ms(ctx, ctypes.byref(e_ms)))
        barrier()
        e_ms = e_ms.value
        if world > 1:
            t = torch.tensor([e_ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms = float(t[0])
        e2e = {"value": N * world * e_steps / (e_ms * 1e-3), "unit": "vectors/s", "h2d_bytes_per_step": N * D * 4,
               "d2h_bytes_per_step": N * 4 + 8, "ms_per_step": e_ms / e_steps,
               "api": "vq_host_step_async (2 lanes; pinned host z -> loss, perplexity, indices on the host)",
               "loss": float(res[0]["loss"][0])}
        lib.vq_host_ctx_destroy(ctx)
    sampler.stop()

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only) ------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        probe = cpu_reference_run(B, D, T, K, 1, 1)
        reps = max(3, min(50, int(args.cpu_seconds / max(probe["ms_per_step"] * 1e-3, 1e-3))))
        r = cpu_reference_run(B, D, T, K, reps, 1)
        cpu = {"value": r["value"], "unit": "vectors/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}

    if rank == 0:
        line = {
            "metric": "quantized vectors/sec (VQ fwd+bwd)", "value": value, "unit": "vectors/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": ("f32" if args.exact else ("f32 (tcgen05 3xTF32 contraction, fp32 accumulate)" if not screen_used else
                      "f32 (tcgen05 TF32 screening pass + exact fp32 refine of the candidates: indices bit-exact vs the fp32 oracle)")),
            "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "B": B, "D": D, "T": T, "K": K, "rows_per_gpu": N, "beta": BETA,
                       "encodings": "dense one-hot emitted" if emit_onehot else "indices only",
                       "path": "exact CUDA-core" if args.exact else ("tcgen05" if lib.vq_forward_uses_tensor_path(N, K, D, fwd_flags) else "exact CUDA-core"),
                       "l2": f"inputs rotate over {nbuf} z + {nbuf} g buffers ({2 * nbuf * N * D * 4 >> 20} MiB > 126 MiB L2)",
                       "parallelism": f"dp{world}: rows sharded, one all-reduce of [dE|hist|sse] per step: {collective}" if world > 1 else "single GPU"},
            "clocks": sampler.summary(),
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "kernels": kern,
            "loss": loss_val, "perplexity": perp_val,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
