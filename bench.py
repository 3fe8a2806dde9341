#!/usr/bin/env python
"""bench.py -- quantized vectors/s for the VectorQuantizer hot path (forward + backward).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

One "step" = one pass of the hot path (vector_quantizer.py:29-58 and its autograd) over one batch of synthetic
latents of the shape `_pre_vq_conv` hands to the quantizer:
    vq_step_forward   codebook norms + distances + argmin + one-hot + gather + losses + perplexity
    vq_step_backward  dz and dE
Default workload = BASELINE.json configs[1]: RIR VQ-VAE quantizer from train_rir.py defaults at batch 256 ->
z (256, 64, 201), N = 51 456 rows, K = 1024, D = 64, beta = 0.25, dense one-hot `encodings` emitted (the reference
always returns it).  N = 1: eager launches chained by programmatic dependent launch; N > 1: replayed from CUDA graphs.

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline          dominant kernel: algorithmic bytes|flops per launch / CUDA-event time per launch
  cpu_baseline      the reference's CPU implementation timed on this box's host cores (bounded sample)
  e2e               same metric through the host-buffer C ABI with the reference's full deliverable copied back
                    (quantized, dz, dE, indices, loss, perplexity); e2e_lean = indices + loss + perplexity only
  kernels           per-kernel share of the step (CUDA events on the launching stream, eager launches)
  sweep             configs[3] corner points (N = 1M rows, indices only): fwd / bwd time, fraction of the roofline
  peaks             HBM and tensor peaks used (TF32 measured here with cuBLAS: burst and sustained)
  module            the drop-in nn.Module on the same workload (eager and inside a CUDA graph)
  collective_check  N > 1: the NVLink exchange against NCCL on the same payload, and the step's dE against an
                    NCCL all-reduce of the local gradients
Under torchrun (N > 1) every rank runs the same per-rank workload on its own rows (weak scaling); the only exchange
is ONE sum all-reduce per step of the packed [dE | usage histogram | squared error] over NVLink peer memory
(vq_dp_allreduce: device-side sequence numbers, so the whole step replays from a CUDA graph; --nccl for NCCL).
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
    "loc16": (16, 64, 201, 1024, "configs[4] RIR side: train_location.py quantizer, batch 16"),
}
SWEEP_K = (512, 1024, 2048, 4096, 8192)
SWEEP_D = (64, 128, 256)
SWEEP_N = 1 << 20
for _k in SWEEP_K:
    for _d in SWEEP_D:
        WORKLOADS[f"sweep_k{_k}_d{_d}"] = (1024, _d, 1024, _k, f"configs[3]: N=1M rows, K={_k}, D={_d}")
SWEEP_CORNERS = ((512, 64), (1024, 64), (4096, 128), (8192, 256))
BETA = 0.25
L2_BYTES = 126 * 1024 * 1024
KERNEL_NAMES = ["prepare_codebook", "forward", "argmin_exact", "rows", "backward", "finalize", "onehot", "exchange"]
METRIC = "quantized vectors/sec (VQ fwd+bwd)"
REFERENCE_ROOTS = ("/root/reference", os.path.join(ROOT, "baseline", "_ref"))


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d["bf16_tflops"]),
                    bf16_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_sustained=1400.0, source="fallback")


def load_traffic():
    """DRAM bytes per launch of each kernel from the committed `ncu --set full` capture (profiles/traffic.json)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {}


def make_config(name, world):
    """The workload description both arms print (identical dictionaries, so the driver can pair them)."""
    B, D, T, K, desc = WORKLOADS[name]
    return {"workload": f"{name}: {desc}", "B": B, "D": D, "T": T, "K": K, "rows_per_gpu": B * T, "beta": BETA,
            "encodings": "indices only" if name.startswith("sweep") else "dense one-hot",
            "parallelism": f"dp{world}" if world > 1 else "single GPU"}


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
# reference arm: the reference's own CPU implementation of the path.  Where the reference tree is present
# (/root/reference in the authoring container, or baseline/_ref) the UNMODIFIED class
# acoustic_locating_vq_vae.vq_vae.vector_quantizer.VectorQuantizer is imported and timed (kind "reference");
# it is a Python module and does not travel to the GPU box, so there the oracle port is timed instead
# (oracle/vq_oracle.py: the same aten calls as vector_quantizer.py:29-58, asserted bit-identical to the class
# by tests/test_oracle.py) -- kind "port".
# ------------------------------------------------------------------------------------------------------
def import_reference_class():
    for root in REFERENCE_ROOTS:
        if os.path.isdir(os.path.join(root, "src", "acoustic_locating_vq_vae")):
            for p in (root, os.path.join(root, "src")):
                if p not in sys.path:
                    sys.path.insert(0, p)
            try:
                from acoustic_locating_vq_vae.vq_vae.vector_quantizer import VectorQuantizer
                return VectorQuantizer, root
            except Exception:
                continue
    return None, None


def cpu_reference_run(name, steps, warmup, budget_s=150.0):
    """`steps` timed steps of forward + `(loss + quantized.sum()).backward()` on all host threads.  The batch is only cut
    (from the front) when the whole run would exceed budget_s; the sample string and rows_per_step say what ran."""
    import torch
    B, D, T, K, _ = WORKLOADS[name]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    E = torch.randn(K, D)
    z = torch.randn(B, D, T)
    cls, root = import_reference_class()
    if cls is not None:
        vq = cls(K, D, BETA)
        vq._embedding.weight.data.copy_(E)
        kind = "reference"

        def run(zz):
            vq._embedding.weight.grad = None
            zz = zz.detach().requires_grad_(True)
            loss, q, perp, enc = vq(zz)
            (loss + q.sum()).backward()
            return float(loss.detach())
    else:
        from oracle import vq_oracle
        kind = "port"

        def run(zz):
            return float(vq_oracle.forward_backward_dense(zz, E, BETA).loss)
    zp = z[:max(1, B // 8)].contiguous()
    run(zp)                                       # cold call (thread pool, allocator): not representative
    t0 = time.perf_counter()
    run(zp)
    t_probe = (time.perf_counter() - t0) * 8
    b_eff = B
    if (steps + warmup) * t_probe > budget_s:
        b_eff = max(1, int(B * budget_s / ((steps + warmup) * t_probe)))
    zs = z[:b_eff].contiguous()
    for _ in range(warmup):
        run(zs)
    t0 = time.perf_counter()
    for _ in range(steps):
        loss = run(zs)
    dt = time.perf_counter() - t0
    rows = b_eff * T
    what = f"unmodified reference class from {root}" if kind == "reference" else "oracle port of the reference class (reference tree absent)"
    return dict(value=rows * steps / dt, ms_per_step=1e3 * dt / steps, cores=cores, rows_per_step=rows, kind=kind, loss=loss,
                sample=f"{what}; {b_eff} of {B} batch items per step ({rows} rows), {steps} timed steps, "
                       f"torch {torch.__version__} CPU, {cores} threads")


def measure_tf32_peak(torch, dev):
    """cuBLAS TF32 GEMM 8192^3 (the yardstick for the tensor-bound sweep points): best of 10, and back to back for ~1.5 s."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device=dev); b = torch.randn(n, n, device=dev); c = torch.empty(n, n, device=dev)
        for _ in range(3):
            torch.matmul(a, b, out=c)
        best = 1e9
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(10):
            e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        reps = max(20, int(1500.0 / best))
        e0.record()
        for _ in range(reps):
            torch.matmul(a, b, out=c)
        e1.record(); torch.cuda.synchronize()
        sus = e0.elapsed_time(e1) / reps
        fl = 2.0 * n ** 3
        return {"tf32_tflops_burst": round(fl / (best * 1e-3) / 1e12, 1), "tf32_tflops_sustained": round(fl / (sus * 1e-3) / 1e12, 1),
                "how": "torch.matmul fp32 with allow_tf32=True (cuBLAS), 8192^3, best of 10 / back to back for ~1.5 s"}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


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
    ap.add_argument("--no-screen", action="store_true", help="3xTF32 tensor kernels instead of screen + exact refine (VQ_FLAG_NO_SCREEN)")
    ap.add_argument("--graph", action="store_true", help="N = 1: replay the step from CUDA graphs (default there: eager launches chained by programmatic dependent launch)")
    ap.add_argument("--no-graph", action="store_true", help="N > 1: eager launches (default there: CUDA graphs, which keep the ranks' launch jitter out of the exchange)")
    ap.add_argument("--dp-mode", default="inline", choices=["inline", "side", "tail", "deferred"], help="N > 1: how the step's exchange is scheduled (see step())")
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: same as --dp-mode inline")
    ap.add_argument("--strong", action="store_true", help="N > 1: strong scaling -- the workload's batch is split over the ranks (default: every rank runs the whole workload, weak scaling)")
    ap.add_argument("--emulate-dp", action="store_true", help="experiment, N = 1: the data-parallel step structure with a world-of-one exchange context")
    ap.add_argument("--nccl", action="store_true", help="N > 1: NCCL all_reduce of [dE|hist|sse] after the backward instead of the NVLink exchange")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the configs[3] corner points")
    ap.add_argument("--sweep-all", action="store_true", help="all 15 (K, D) sweep points instead of the 4 corners")
    ap.add_argument("--no-module", action="store_true", help="skip the nn.Module timing")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    B, D, T, K, desc = WORKLOADS[args.workload]
    if args.strong and world > 1:          # total work fixed: the workload's batch items are split over the ranks
        if B % world:
            raise SystemExit(f"--strong: batch {B} does not divide over {world} ranks")
        B //= world
    N = B * T
    # configs[3] (the sweep) is indices-only by definition (a dense one-hot would be up to 34 GB)
    emit_onehot = not args.no_onehot and not args.workload.startswith("sweep")
    config = make_config(args.workload, max(world, args.gpus))
    if args.strong and world > 1:
        config.update({"B": B, "rows_per_gpu": N, "parallelism": f"dp{world} (strong scaling: the workload's batch split over the ranks)"})
    if args.no_onehot:
        config["encodings"] = "indices only"

    # -------------------------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        r = cpu_reference_run(args.workload, args.steps, args.warmup)
        if r["rows_per_step"] != N:
            config["rows_per_gpu"] = r["rows_per_step"]      # the budget forced a smaller batch: say so where the driver looks
        line = {
            "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "vectors/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": r["value"], "unit": "vectors/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "vectors/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "loss": r["loss"],
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
    P = lambda t: None if t is None else t.data_ptr()

    def make_state(N, K, D, onehot_on, nbuf, seed, nsets=2):
        """device buffers of one workload; inputs rotate over `nbuf` z / g_q buffers"""
        g = torch.Generator(device=dev)
        g.manual_seed(seed)
        s = dict(N=N, K=K, D=D, nbuf=nbuf)
        torch.manual_seed(0)
        s["E"] = torch.randn(K, D).to(dev)                              # same codebook on every rank
        s["zs"] = [torch.randn(N, D, device=dev, generator=g) for _ in range(nbuf)]
        s["gs"] = [torch.randn(N, D, device=dev, generator=g) for _ in range(nbuf)]
        s["e2"] = torch.empty(K, device=dev); s["ehi"] = torch.empty(K, D, device=dev); s["elo"] = torch.empty(K, D, device=dev)
        s["q"] = torch.empty(N, D, device=dev); s["idx"] = torch.empty(N, dtype=torch.int32, device=dev)
        s["onehot"] = torch.empty(N, K, device=dev) if onehot_on else None
        # the packed step buffer [dE (K*D) | hist (K) | sse] + loss, perplexity: the forward writes the statistics and
        # the backward the gradient straight into it, so data parallel all-reduces it as it stands
        # `nsets` sets, rotating over the steps under data parallelism: the exchange of step i runs on its own stream while
        # the following steps already fill the other sets
        s["packs"] = [torch.zeros(K * D + K + 3, device=dev) for _ in range(nsets)]
        s["reds"] = [torch.zeros(K * D + K + 1, device=dev) for _ in range(nsets)]   # data parallel: the all-reduced packed buffers
        s["packed"], s["reduced"] = s["packs"][0], s["reds"][0]
        s["stats"] = s["packed"][K * D:]                                # [hist (K) | sse | loss | perplexity]
        s["dE"] = s["packed"][:K * D].view(K, D)
        s["dz"] = torch.empty(N, D, device=dev)
        s["g_loss"] = torch.ones((), device=dev)
        s["wsb"] = lib.vq_workspace_bytes(N, K, D, 0)
        s["ws"] = torch.empty(s["wsb"], dtype=torch.uint8, device=dev)
        s["fwd_flags"] = (L.FLAG_ONEHOT if onehot_on else 0) | (L.FLAG_EXACT if args.exact else 0) | (L.FLAG_NO_SCREEN if args.no_screen else 0)
        return s

    nbuf = max(3, int(1.25 * L2_BYTES / (N * D * 4)) + 1)                # rotating inputs: set larger than L2
    S0 = make_state(N, K, D, emit_onehot, nbuf, 1000 + rank, nsets=min(2 * nbuf, 31) if (world > 1 or args.emulate_dp) else 1)
    n_packed = K * D + K + 1
    exch = None
    collective = "none"
    if world > 1 and not args.nccl:
        par = import_module("acoustic_locating_vq-vae_b200.parallel")
        ok, why = True, ""
        try:
            exch = par.PeerExchange(n_packed, dev)
        except Exception as e:          # symmetric memory unavailable on this rank
            ok, why = False, f"{type(e).__name__}"
        if not par.agree(ok):           # every rank takes the same collective
            if exch is not None:
                exch.close()
            exch = None
            collective = f"NCCL all_reduce (NVLink exchange unavailable on some rank{': ' + why if why else ''})"
        else:
            collective = ("[dE | hist | sse] over NVLink peer memory (vq_dp_allreduce, " + ("NVLS multimem.st" if exch.nvls else "P2P stores") + ", "
                          + ("reduce-scatter + all-gather" if world >= 8 else "one step") + ", device-side sequence numbers"
                          + "; schedule: " + {"tail": "in the tail of the backward kernel (vq_step_backward_dp): the last CTAs to finish push / collect / sum, no second launch",
                                             "inline": "after the fused backward, nothing overlaps it",
                                             "deferred": "one step late, between the next step's forward and backward (overlaps that backward)",
                                             "side": "on the context's own low-priority stream, started behind the next step's prepare launch"}["inline" if args.no_overlap else args.dp_mode] + ")")
    elif world > 1:
        collective = "NCCL all_reduce of [dE | hist | sse] after the backward"
    elif args.emulate_dp:
        # experiment (one GPU): the data-parallel step structure with a world-of-one exchange context -- the launch
        # structure (fork / join, graph edges, kernel co-residency) without NVLink or a second rank
        class _Solo:
            nvls = False

            def __init__(self):
                lines = int(lib.vq_dp_recv_lines(1, n_packed))
                self.bufs = [torch.zeros(lines * 4, device=dev) for _ in range(2)]
                arr = [(ctypes.c_void_p * 1)(b.data_ptr()) for b in self.bufs]
                self.ctx = ctypes.c_void_p()
                L.check(lib.vq_dp_create(arr[0], arr[1], None, None, 1, 0, n_packed, 0, ctypes.byref(self.ctx)))

            def status(self, st):
                calls, err = ctypes.c_uint32(0), ctypes.c_uint32(0)
                L.check(lib.vq_dp_status(self.ctx, ctypes.byref(calls), ctypes.byref(err), st))
                return calls.value, err.value

            def close(self):
                if self.ctx is not None:
                    lib.vq_dp_destroy(self.ctx)
                    self.ctx = None
        exch = _Solo()
        collective = "EMULATION: world-of-one exchange context on a single GPU (launch structure only)"
    n_dE_scale = world

    # how the one exchange of a data-parallel step is scheduled (N > 1):
    #   inline    (default) fused backward, then the exchange kernel, chained by programmatic dependent launch
    #   tail      inside the backward kernel (vq_step_backward_dp): the last CTAs to finish ARE the exchange (no second launch;
    #             measured slower: the serial round trips of the exchange stay, and the long-lived CTAs stream slower)
    #   deferred  the exchange of step i sits between forward and backward of step i + 1 (one step late; two buffer sets)
    #   side      the exchange of step i on the context's own low-priority stream, started behind step i + 1's prepare launch,
    #             joined at the end of every graph
    # Measured: nothing overlaps the exchange with the fused forward, which owns every SM's registers
    # and shared memory -- a co-scheduled exchange either holds the forward's CTAs up or starves until the graph's end.
    dp_mode = "none" if exch is None else ("inline" if args.no_overlap else args.dp_mode)
    sched = {"mode": dp_mode}                         # (the per-kernel profile pass runs the exchange in line)
    pending = []                                     # deferred / side: the previous step's (payload, result), exchange not started yet

    def step(s, i, st):
        """one step of the hot path on input buffer i, enqueued on raw stream `st`"""
        N_, K_, D_ = s["N"], s["K"], s["D"]
        z, gq = s["zs"][i % s["nbuf"]], s["gs"][i % s["nbuf"]]
        ns = len(s["packs"])
        par = (i % ns) if sched["mode"] in ("side", "deferred") else 0
        pk = P(s["packs"][par])                      # [dE (K*D) | hist (K) | sse | loss | perplexity]
        red = P(s["reds"][par])
        sp = pk + 4 * K_ * D_
        if sched["mode"] == "side":
            # the exchange that last used this set of buffers (step i - ns) is complete (a graph of <= ns steps never waits
            # here: its only join is the final one)
            L.check(lib.vq_dp_wait(exch.ctx, max(ns - 2, 0), st))
            # codebook preparation (also zeroes the dE accumulator) ...
            need_elo = args.no_screen or not (K_ % 256 == 0 and D_ in (32, 64, 96, 128, 192, 256))     # as vq_step_forward decides
            L.check(lib.vq_prepare_step(P(s["E"]), K_, D_, P(s["e2"]), None if args.exact else P(s["ehi"]), P(s["elo"]) if (need_elo and not args.exact) else None,
                                        sp, P(s["ws"]), s["wsb"], pk, st))
            # ... then the PREVIOUS step's exchange is started, behind this step's prepare launch: the fused forward owns
            # every SM's registers and shared memory, so an exchange that gets its CTAs onto the SMs first would hold the
            # forward up; started here (on a lower-priority stream) it yields to the forward, which is already being placed
            if pending:
                L.check(lib.vq_dp_allreduce_start(exch.ctx, pending[0], pending[1], st))
                del pending[:]
            L.check(lib.vq_forward(P(z), P(s["E"]), P(s["e2"]), P(s["ehi"]), P(s["elo"]), N_, K_, D_, BETA, s["fwd_flags"] | L.FLAG_STATE_READY, P(s["q"]), P(s["idx"]),
                                   P(s["onehot"]), sp, sp + 4 * K_, sp + 4 * (K_ + 1), sp + 4 * (K_ + 2), P(s["ws"]), s["wsb"], st))
        else:
            # the prepare launch also zeroes the dE accumulator: no memset node between forward and backward (PDL chain intact)
            L.check(lib.vq_step_forward(P(z), P(s["E"]), N_, K_, D_, BETA, s["fwd_flags"], P(s["e2"]), P(s["ehi"]), P(s["elo"]), pk, P(s["q"]), P(s["idx"]),
                                        P(s["onehot"]), sp, sp + 4 * K_, sp + 4 * (K_ + 1), sp + 4 * (K_ + 2), P(s["ws"]), s["wsb"], st))
        n_dE = N_ * n_dE_scale
        if sched["mode"] == "tail":           # data parallel: ONE sum all-reduce of [dE | hist | sse], in the backward kernel's tail
            L.check(lib.vq_step_backward_dp(P(gq), P(s["g_loss"]), P(z), P(s["E"]), P(s["idx"]), N_, N_, n_dE, K_, D_, BETA, L.FLAG_TRAIN_VQ,
                                            P(s["dz"]), pk, exch.ctx, red, P(s["ws"]), s["wsb"], s["fwd_flags"], st))
            return
        if sched["mode"] == "deferred" and pending:          # the previous step's exchange: overlaps this step's backward
            L.check(lib.vq_dp_allreduce(exch.ctx, pending[0], pending[1], st))
            del pending[:]
        # right behind the forward: starts on the workspace's ready word and overlaps the forward's statistics tail
        L.check(lib.vq_step_backward(P(gq), P(s["g_loss"]), P(z), P(s["E"]), P(s["idx"]), N_, N_, n_dE, K_, D_, BETA, L.FLAG_TRAIN_VQ,
                                     P(s["dz"]), pk, P(s["ws"]), s["wsb"], s["fwd_flags"], st))
        if sched["mode"] in ("side", "deferred"):
            pending[:] = [pk, red]                                                        # started by the next step (or the join)
        elif sched["mode"] == "inline":
            L.check(lib.vq_dp_allreduce(exch.ctx, pk, red, st))
        elif world > 1:
            dist.all_reduce(s["packs"][par][:K_ * D_ + K_ + 1])

    def join(st):
        """every exchange is started and complete before `st` goes on (end of a run, end of a capture)"""
        if pending:
            if sched["mode"] == "side":
                L.check(lib.vq_dp_allreduce_start(exch.ctx, pending[0], pending[1], st))
            else:
                L.check(lib.vq_dp_allreduce(exch.ctx, pending[0], pending[1], st))
            del pending[:]
        if sched["mode"] == "side":
            L.check(lib.vq_dp_wait(exch.ctx, 0, st))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # CUDA graphs take the host out of the loop (the exchange carries its sequence number on the device, so the data-parallel
    # step replays too).  Programmatic dependent launch does not cross graph boundaries, so one graph holds a whole round over
    # the input buffers (nbuf steps); single-step graphs cover the remainder.  One GPU: eager by default -- the step is
    # GPU-bound, the host stays ahead and the PDL chain is never broken (68.8 vs 72.2 us).  Data parallel: graphs by default --
    # every rank's Python launch jitter otherwise turns into waiting inside the exchange (2 GPUs: 83.0 - 83.5 us replayed,
    # 83.2 - 90.7 us eager).
    use_graph = (args.graph or ((world > 1 or args.emulate_dp) and not args.no_graph)) and not (world > 1 and exch is None)   # NCCL stays eager
    cur = torch.cuda.current_stream()

    def build_runner(s):
        """returns run_steps(k): enqueues k consecutive steps starting at input buffer 0"""
        nb = s["nbuf"]
        if not use_graph:
            def run_eager(k):
                for i in range(k):
                    step(s, i, cur.cuda_stream)
                join(cur.cuda_stream)
            run_eager.prime = lambda k: None
            return run_eager, {}
        for i in range(min(3, nb)):          # warm every code path before capture (lazy attribute setting, descriptors)
            step(s, i, cur.cuda_stream)
        join(cur.cuda_stream)
        torch.cuda.synchronize()
        side = torch.cuda.Stream(device=dev, priority=-1)     # the step's kernels outrank the exchange's (captured with the nodes)
        graphs = {}

        def capture(n):
            """one graph holding steps 0 .. n-1 (programmatic edges inside, the exchange's fork / join edges under data parallelism)"""
            torch.cuda.synchronize()
            side.wait_stream(cur)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side):
                    for i in range(n):
                        step(s, i, torch.cuda.current_stream().cuda_stream)
                    join(torch.cuda.current_stream().cuda_stream)     # a forked stream rejoins before the capture ends
            cur.wait_stream(side)
            torch.cuda.synchronize()
            graphs[n] = g

        # steps per graph: one round over the input buffers; data parallel: as many steps as there are payload / result sets
        # (two rounds), so that no join but the final one sits inside a graph and a 20-step run is ONE launch
        chunk = len(s["packs"]) if exch is not None else nb

        def plan(k):
            return [chunk] * (k // chunk) + ([k % chunk] if k % chunk else [])

        def run_graphs(k):
            for n in plan(k):
                if n not in graphs:          # only outside timed regions: the bench primes the sizes it needs
                    capture(n)
                graphs[n].replay()
        def prime(k):
            """capture the graphs a k-step run needs and replay each once, untimed: the first launch of an instantiated graph
            uploads it to the device (tens of microseconds for 100 kernel nodes), which a 20-step region would count"""
            for n in set(plan(k)):
                if n not in graphs:
                    capture(n)
                    graphs[n].replay()
            torch.cuda.synchronize()
        run_graphs.prime = prime
        run_graphs.graphs = graphs
        return run_graphs, graphs

    run0, n_graphs = build_runner(S0)
    l0 = lib.vq_launch_count()
    step(S0, 0, cur.cuda_stream)
    join(cur.cuda_stream)
    launches_per_step = lib.vq_launch_count() - l0
    barrier()

    sampler = ClockSampler(local_rank)
    # ---- device-resident throughput ------------------------------------------------------------------
    run0.prime(args.warmup)
    run0.prime(args.steps)
    run0(args.warmup)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if exch is not None and world > 1:
        # the host barrier leaves the ranks tens of microseconds apart, which a 20-step region (1.4 ms) would count as
        # exchange time: one untimed exchange lines the DEVICE timelines up (nobody leaves it before everybody has entered)
        L.check(lib.vq_dp_allreduce(exch.ctx, P(S0["packs"][0]), P(S0["reds"][0]), cur.cuda_stream))
    ev0.record()
    run0(args.steps)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    value = N * world * args.steps / (ms * 1e-3)
    loss_val, perp_val = [float(x) for x in S0["stats"][K + 1:].tolist()]
    dp_status = None
    if exch is not None:
        calls, err = exch.status(cur.cuda_stream)
        dp_status = {"calls": calls, "error_word": err}

    # ---- per-kernel timing (CUDA events on the launching stream, eager launches) for the roofline ---------
    def profile(s, reps):
        keep_mode = sched["mode"]
        if keep_mode in ("side", "deferred"):        # time the exchange where it has the SMs to itself, not while it queues behind the forward
            sched["mode"] = "inline"
        lib.vq_profile_enable(1)
        for i in range(reps):
            step(s, i, cur.cuda_stream)
        join(cur.cuda_stream)
        barrier()
        sched["mode"] = keep_mode
        kern = {}
        for kid, name in enumerate(KERNEL_NAMES):
            tot, cnt = ctypes.c_double(0), ctypes.c_int64(0)
            L.check(lib.vq_profile_read(kid, ctypes.byref(tot), ctypes.byref(cnt)))
            if cnt.value:
                kern[name] = {"launches": cnt.value, "avg_us": 1e3 * tot.value / cnt.value}
        lib.vq_profile_enable(0)
        return kern

    kern = profile(S0, args.steps)
    tot_us = sum(k["avg_us"] * k["launches"] for k in kern.values()) / max(args.steps, 1)
    for k in kern.values():
        k["share"] = round(k["avg_us"] * k["launches"] / args.steps / tot_us, 4)
        k["avg_us"] = round(k["avg_us"], 3)
    dom = max(kern, key=lambda n: kern[n]["avg_us"] * kern[n]["launches"])
    traffic = load_traffic()
    fwd_bytes = lambda N_, K_, D_, oh: 4.0 * (2 * N_ * D_ + N_ + K_ * D_ + K_ + (N_ * K_ if oh else 0))
    bwd_bytes = lambda N_, K_, D_: 4.0 * (3 * N_ * D_ + N_ + 2 * K_ * D_)
    # algorithmic work per launch: SURVEY.md section 8(d) per-row figures x rows per launch
    alg = {
        "forward": {"flops": 2.0 * N * K * D, "bytes": fwd_bytes(N, K, D, emit_onehot)},
        "argmin_exact": {"flops": 2.0 * N * K * D, "bytes": 4.0 * (N * D + N + K * D)},
        "rows": {"flops": 0.0, "bytes": fwd_bytes(N, K, D, emit_onehot)},
        "backward": {"flops": 0.0, "bytes": bwd_bytes(N, K, D)},
        "prepare_codebook": {"flops": 0.0, "bytes": 4.0 * (3 * K * D + K)},
        "exchange": {"flops": 0.0, "bytes": 4.0 * 2 * n_packed},
    }
    work = alg.get(dom, {"flops": 0.0, "bytes": 0.0})
    dur_s = kern[dom]["avg_us"] * 1e-6
    # The TF32 peak (cuBLAS 8192^3, back to back for ~1.5 s) is measured AFTER the sweep: the sustained GEMM drives the GPU
    # into its power cap, and sweep points timed right behind it ran up to 1.7x slower on some boxes.  Until then the
    # bound decision uses bf16 / 2; a tensor-bound roofline is re-based on the measured burst peak below.
    tf32 = None
    tf32_burst = peaks["bf16_tflops"] / 2.0
    tf32_sus = peaks["bf16_sustained"] / 2.0
    tf32_src = peaks["source"] + " bf16 / 2"
    t_tensor = work["flops"] / (tf32_burst * 1e12)
    t_hbm = work["bytes"] / (peaks["hbm_gbs"] * 1e9)
    if t_tensor >= t_hbm:                                          # whichever roofline binds this launch
        bound, peak, achieved, unit, per_launch = "tensor", tf32_burst, work["flops"] / dur_s / 1e12, "TFLOP/s", work["flops"]
    else:
        bound, peak, achieved, unit, per_launch = "hbm", peaks["hbm_gbs"], work["bytes"] / dur_s / 1e9, "GB/s", work["bytes"]
    tensor_path = bool(lib.vq_forward_uses_tensor_path(N, K, D, S0["fwd_flags"]))
    screen_used = tensor_path and not args.no_screen and K % 256 == 0 and D in (32, 64, 96, 128, 192, 256) and os.environ.get("B200VQ_SCREEN", "1")[:1] != "0"
    kname = {"forward": "vq_screen_kernel (fused forward: TF32 screen + exact refine, row epilogue, one-hot)" if screen_used
                        else "fused forward (argmin_tc2: 3xTF32 + row epilogue)"}.get(dom, dom)
    tr = traffic.get(args.workload, {})
    roofline = {"kernel": kname, "bound": bound, "achieved": round(achieved, 2), "peak": round(peak, 1), "unit": unit,
                "frac": round(achieved / peak, 4), "traffic": tr.get(dom),
                "traffic_source": tr.get("source", "profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one committed ncu --set full capture (not measured in this run)"),
                "peak_source": (tf32_src if bound == "tensor" else peaks["source"] + " copy bandwidth"),
                "per_launch": per_launch, "avg_us": kern[dom]["avg_us"],
                "step": {"bytes": fwd_bytes(N, K, D, emit_onehot) + bwd_bytes(N, K, D), "us": round(1e3 * ms / args.steps, 3),
                         "frac_of_hbm": round((fwd_bytes(N, K, D, emit_onehot) + bwd_bytes(N, K, D)) / (1e-3 * ms / args.steps) / 1e9 / peaks["hbm_gbs"], 4)},
                "other_roof": {"tensor_tflops": round(work["flops"] / dur_s / 1e12, 2), "hbm_gbs": round(work["bytes"] / dur_s / 1e9, 1)}}

    # ---- data parallel: the exchange against NCCL, outside every timed region ---------------------------
    coll_check = None
    if world > 1 and exch is not None:
        torch.manual_seed(77 + rank)
        pay = torch.randn(n_packed, device=dev)
        out = torch.empty_like(pay)
        exch.allreduce(pay, out, cur.cuda_stream)
        ref = pay.clone()
        dist.all_reduce(ref)
        d1 = float((out - ref).abs().max())
        # the step's gradient: DP step vs NCCL all-reduce of the local gradients (same rows, same scale)
        step(S0, 0, cur.cuda_stream)
        join(cur.cuda_stream)
        torch.cuda.synchronize()
        dE_dp = S0["reduced"][:K * D].view(K, D).clone()
        stats_dp = S0["reduced"][K * D:].clone()
        loc = torch.zeros(K, D, device=dev)
        L.check(lib.vq_backward(P(S0["gs"][0]), P(S0["g_loss"]), P(S0["zs"][0]), P(S0["E"]), P(S0["idx"]), N, N, N * world, K, D, BETA,
                                L.FLAG_TRAIN_VQ | L.FLAG_ZERO_DE | L.FLAG_BWD_FLAT, P(S0["dz"]), P(loc), cur.cuda_stream))
        dist.all_reduce(loc)
        st_loc = S0["stats"][:K + 1].clone()
        dist.all_reduce(st_loc)
        d2 = float((dE_dp - loc).abs().max())
        d3 = float((stats_dp - st_loc).abs().max())
        chk = torch.tensor([float(out.double().sum()), float(dE_dp.double().sum())], dtype=torch.float64, device=dev)
        gathered = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(gathered, chk)
        same = all(torch.equal(gathered[0], g_) for g_ in gathered)
        coll_check = {"max_abs_diff": d1, "vs": "dist.all_reduce (NCCL) on the same random payload", "payload_floats": n_packed,
                      "step_dE_max_abs_diff": d2, "step_dE_scale": float(loc.abs().max()), "step_stats_max_abs_diff": d3,
                      "checksums_identical_on_all_ranks": bool(same), "status": dp_status}

    # ---- end to end through the host-buffer C ABI ---------------------------------------------------------
    def run_e2e(full):
        ctx = ctypes.c_void_p()
        L.check(lib.vq_host_ctx_create(N, K, D, ctypes.byref(ctx)))
        E_host = S0["E"].cpu().contiguous()
        L.check(lib.vq_host_set_codebook(ctx, E_host.data_ptr()))
        nhost = 4
        z_host = [torch.randn(N, D).pin_memory() for _ in range(nhost)]
        res = [dict(loss=torch.zeros(1).pin_memory(), perp=torch.zeros(1).pin_memory(), idx=torch.zeros(N, dtype=torch.int32).pin_memory(),
                    q=torch.zeros(N, D).pin_memory() if full else None, dz=torch.zeros(N, D).pin_memory() if full else None,
                    dE=torch.zeros(K, D).pin_memory() if full else None) for _ in range(2)]
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
            L.check(lib.vq_host_step_async(ctx, lane, z_host[i % nhost].data_ptr(), None, N, N * world, BETA,
                                           L.FLAG_TRAIN_VQ | (L.FLAG_EXACT if args.exact else 0),
                                           r["loss"].data_ptr(), r["perp"].data_ptr(), r["idx"].data_ptr(), P(r["q"]), P(r["dz"]),
                                           None if (world > 1 or not full) else P(r["dE"])))
            if world > 1:                                 # dE crosses the ranks before it goes home
                s_, t_ = lane_t[lane]
                with torch.cuda.stream(s_):
                    dist.all_reduce(t_)
                    if full:
                        r["dE"].view(-1).copy_(t_, non_blocking=True)

        for i in range(args.warmup):
            host_step(i)
        L.check(lib.vq_host_wait(ctx, 0)); L.check(lib.vq_host_wait(ctx, 1))
        barrier()
        L.check(lib.vq_host_timer_start(ctx))
        for i in range(args.steps):
            host_step(i)
        e_ms = ctypes.c_float(0)
        L.check(lib.vq_host_timer_stop_ms(ctx, ctypes.byref(e_ms)))
        barrier()
        e_ms = e_ms.value
        if world > 1:
            t = torch.tensor([e_ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms = float(t[0])
        d2h = N * 4 + 8 + ((2 * N * D + K * D) * 4 if full else 0)
        out = {"value": N * world * args.steps / (e_ms * 1e-3), "unit": "vectors/s", "h2d_bytes_per_step": N * D * 4, "d2h_bytes_per_step": d2h,
               "ms_per_step": e_ms / args.steps, "loss": float(res[0]["loss"][0]),
               "api": "vq_host_step_async (2 lanes; pinned host z in; " + ("quantized, dz, dE, indices, loss, perplexity" if full else "indices, loss, perplexity")
                      + " back on the host)"}
        if full:
            out["encodings"] = "indices (N int32); the dense (N, K) one-hot is rebuilt by the caller (one store per row), not shipped over PCIe"
        # tensors that were used on the lanes' streams (torch remembers the stream of every non-blocking copy and all-reduce)
        # go before the streams do
        res.clear(); lane_t.clear(); z_host.clear()
        torch.cuda.synchronize()
        lib.vq_host_ctx_destroy(ctx)
        return out

    e2e = e2e_lean = None
    if not args.skip_e2e:
        e2e = run_e2e(True)
        e2e_lean = run_e2e(False)
    sampler.stop()

    # ---- the drop-in nn.Module on the same workload (north star: the module is the deliverable) ------------------
    module = None
    if world == 1 and not args.no_module:
        torch.cuda.empty_cache()
        vq = b200vq.VectorQuantizer(K, D, BETA, return_encodings=emit_onehot).to(dev)
        vq._embedding.weight.data.copy_(S0["E"])
        zin = S0["zs"][0].view(B, D, T).clone().requires_grad_(True)      # what _pre_vq_conv hands over: contiguous (B, D, T)
        gq = torch.ones(B, D, T, device=dev)

        def mstep():
            vq._embedding.weight.grad = None
            zin.grad = None
            loss, q, perp, enc = vq(zin)
            torch.autograd.backward([loss, q], [None, gq])
        for _ in range(10):
            mstep()
        torch.cuda.synchronize()
        reps = 50
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            mstep()
        b.record(); torch.cuda.synchronize()
        module = {"eager_us_per_step": round(a.elapsed_time(b) / reps * 1e3, 1), "what": "b200vq.VectorQuantizer forward + autograd backward, "
                  + ("dense one-hot returned" if emit_onehot else "return_encodings=False")}
        try:
            gm = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                mstep()
                with torch.cuda.graph(gm, stream=side):
                    mstep()
            cur.wait_stream(side)
            torch.cuda.synchronize()
            a.record()
            for _ in range(reps):
                gm.replay()
            b.record(); torch.cuda.synchronize()
            module["graph_us_per_step"] = round(a.elapsed_time(b) / reps * 1e3, 1)
        except Exception as e:
            module["graph_us_per_step"] = None
            module["graph_error"] = f"{type(e).__name__}: {e}"[:200]

    # ---- configs[3]: sweep corner points (N = 1M rows, indices only), single GPU ----------------------------
    sweep = None
    if world == 1 and not args.no_sweep and not args.workload.startswith("sweep"):
        raw = []
        pts = [(k, d) for k in SWEEP_K for d in SWEEP_D] if args.sweep_all else SWEEP_CORNERS
        del S0["zs"][1:], S0["gs"][1:]
        for (k_, d_) in pts:
            torch.cuda.empty_cache()
            s = make_state(SWEEP_N, k_, d_, False, 1 if d_ >= 128 else 2, 5)
            for i in range(3):
                step(s, i, cur.cuda_stream)
            torch.cuda.synchronize()
            reps = 5
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(reps):
                step(s, i, cur.cuda_stream)
            b.record(); torch.cuda.synchronize()
            t_step = a.elapsed_time(b) / reps * 1e3
            kk = profile(s, reps)
            # forward = prepare + fused forward (+ the rows kernel where the launcher splits the row epilogue off: small K)
            t_f = (kk.get("forward", kk.get("argmin_exact", {"avg_us": float("nan")}))["avg_us"] + kk.get("prepare_codebook", {"avg_us": 0.0})["avg_us"]
                   + kk.get("rows", {"avg_us": 0.0})["avg_us"])
            t_b = kk["backward"]["avg_us"]
            raw.append((k_, d_, t_f, t_b, t_step, ["flat", "", "private", "replicated"][lib.vq_backward_path(SWEEP_N, k_, d_, 0)],
                        sorted(n for n in kk if n not in ("backward", "exchange"))))
            del s
        torch.cuda.empty_cache()
        tf32 = measure_tf32_peak(torch, dev)
        tf32_burst, tf32_sus, tf32_src = tf32["tf32_tflops_burst"], tf32["tf32_tflops_sustained"], "measured here (cuBLAS TF32 8192^3), after the sweep"
        sweep = {"n_rows": SWEEP_N, "encodings": "indices only", "tf32_peak_tflops": {"burst": tf32_burst, "sustained": tf32_sus, "source": tf32_src},
                 "note": "frac_fwd = algorithmic 2NKD / t_fwd over the sustained TF32 peak; frac = T_roof / (t_fwd + t_bwd), "
                         "T_roof = max(2NKD / P_tf32, bytes_fwd / BW) + bytes_bwd / BW (SURVEY.md 8d)", "points": []}
        for (k_, d_, t_f, t_b, t_step, bpath, fkern) in raw:
            fl = 2.0 * SWEEP_N * k_ * d_
            bf, bb = fwd_bytes(SWEEP_N, k_, d_, False), bwd_bytes(SWEEP_N, k_, d_)
            t_roof = max(fl / (tf32_sus * 1e12), bf / (peaks["hbm_gbs"] * 1e9)) + bb / (peaks["hbm_gbs"] * 1e9)
            sweep["points"].append({"K": k_, "D": d_, "fwd_us": round(t_f, 1), "bwd_us": round(t_b, 1), "step_us": round(t_step, 1),
                                    "fwd_tflops": round(fl / (t_f * 1e-6) / 1e12, 1), "frac_fwd": round(fl / (t_f * 1e-6) / 1e12 / tf32_sus, 3),
                                    "bwd_frac_hbm": round(bb / (t_b * 1e-6) / 1e9 / peaks["hbm_gbs"], 3),
                                    "frac": round(t_roof * 1e6 / (t_f + t_b), 3), "vectors_per_s": round(SWEEP_N / (t_step * 1e-6)),
                                    "backward_path": bpath, "forward_kernels": fkern})
    elif world == 1 and not args.no_sweep:      # a sweep point is the main workload: its roofline wants the measured peak
        tf32 = measure_tf32_peak(torch, dev)
        tf32_burst, tf32_sus, tf32_src = tf32["tf32_tflops_burst"], tf32["tf32_tflops_sustained"], "measured here (cuBLAS TF32 8192^3), after the timed regions"
    if tf32 is not None and roofline["bound"] == "tensor":
        roofline.update({"peak": round(tf32_burst, 1), "frac": round(roofline["achieved"] / tf32_burst, 4), "peak_source": tf32_src})

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only) ------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        probe = cpu_reference_run(args.workload, 1, 1)
        reps = max(3, min(50, int(args.cpu_seconds / max(probe["ms_per_step"] * 1e-3, 1e-3))))
        r = cpu_reference_run(args.workload, reps, 1)
        cpu = {"value": r["value"], "unit": "vectors/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "vectors/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if (args.strong and world > 1) else "weak", "vs_baseline": None,
            "dtype": ("f32" if args.exact or not tensor_path else ("f32 (tcgen05 3xTF32 contraction, fp32 accumulate)" if not screen_used else
                      "f32 (tcgen05 TF32 screening pass + exact fp32 refine of the candidates: indices bit-exact vs the fp32 oracle)")),
            "data": "synthetic", "config": config,
            "notes": {"path": "tcgen05" if tensor_path else "exact CUDA-core",
                      "launch": (f"CUDA graphs ({len(S0['packs']) if exch is not None else nbuf} steps per graph + one for the remaining steps; inputs rotate over {nbuf} buffers), {launches_per_step} kernels per step"
                                 if use_graph else f"eager launches chained by programmatic dependent launch, {launches_per_step} kernels per step"),
                      "l2": f"inputs rotate over {nbuf} z + {nbuf} g buffers ({2 * nbuf * N * D * 4 >> 20} MiB > 126 MiB L2)",
                      "collective": collective,
                      "kernel_timing": "kernels.*: CUDA events around every launch on its own stream (eager, profile pass). The backward is launched early (programmatic dependent "
                                       "launch) and starts on the forward's ready word, so its figure includes that wait; stand-alone it takes 11 us back to back "
                                       "(14.5 - 15.6 us cold and serialised under ncu: profiles/r2_launches_rir256_s1.csv)"},
            "clocks": sampler.summary(),
            "e2e": e2e, "e2e_lean": e2e_lean, "gpu_launches": int(launches_per_step * args.steps), "roofline": roofline, "cpu_baseline": cpu,
            "kernels": kern, "peaks": {"hbm_gbs": peaks["hbm_gbs"], "hbm_source": peaks["source"], "tf32": tf32}, "sweep": sweep, "module": module,
            "collective_check": coll_check, "loss": loss_val, "perplexity": perp_val,
        }
        print(json.dumps(line), flush=True)
    if exch is not None:
        exch.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
