"""GPU tests of the backward strategies -- vq_backward's flat kernel (one red.global.add per
element) and private kernel (per-CTA copy of dE in shared memory) -- and of the data-parallel exchange.  Both kernels
must give the oracle's gradients (autograd of vector_quantizer.py:46-54, restated in oracle/vq_oracle.c) within the
north-star tolerance; dz is the same arithmetic in both and must be BIT-identical.  Edge cases: ragged N, K not a
multiple of anything, a single hot code that owns every row, accumulate-vs-overwrite semantics of VQ_FLAG_ZERO_DE,
dz == NULL, global row counts, codes outside [0, K).  The exchange runs with `world` ranks emulated on one GPU.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
BETA = 0.25
TRAIN, ZERO_DE, FLAT, PRIVATE = 1 << 1, 1 << 5, 1 << 12, 1 << 14


def _dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _backward(lib, g, gl, z, E, idx, flags, n_dE=None, dE_init=None, want_dz=True):
    dev = z.device
    N, D = z.shape
    K = E.shape[0]
    dz = torch.full((N, D), 7.0, device=dev) if want_dz else None
    dE = torch.zeros(K, D, device=dev) if dE_init is None else dE_init.clone()
    glt = torch.tensor(gl, dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.vq_backward(None if g is None else g.data_ptr(), glt.data_ptr(), z.data_ptr(), E.data_ptr(), idx.data_ptr(), N, N,
                         N if n_dE is None else n_dE, K, D, BETA, flags, None if dz is None else dz.data_ptr(), dE.data_ptr(), st)
    assert rc == 0, lib.vq_last_error()
    torch.cuda.synchronize()
    return dz, dE


def _oracle(g, gl, z, E, idx):
    from oracle import c_oracle
    return c_oracle.backward(g.cpu().numpy(), gl, z.cpu().numpy(), E.cpu().numpy(), idx.cpu().numpy().astype(np.int32), BETA, True)


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def _make(N, D, K, seed, hot=None):
    dev = _dev()
    g0 = torch.Generator(device="cpu").manual_seed(seed)
    E = torch.randn(K, D, generator=g0).to(dev)
    z = torch.randn(N, D, generator=g0).to(dev)
    g = torch.randn(N, D, generator=g0).to(dev)
    if hot is None:
        # a skewed but realistic assignment: nearest code under the reference's distance (cheap torch ops; the index
        # source does not matter for the backward, any idx in [0, K) is a valid input)
        d = (z * z).sum(1, keepdim=True) + (E * E).sum(1) - 2 * z @ E.t()
        idx = d.argmin(1).int()
    else:
        idx = torch.full((N,), hot, dtype=torch.int32, device=dev)
    return E, z, g, idx


SHAPES = [
    # N, D, K
    (51456, 64, 1024),      # BASELINE configs[1]
    (16000, 128, 1024),     # configs[0]
    (3216, 64, 1024),       # configs[4] RIR side
    (4099, 64, 200),        # ragged N, K not a multiple of 128
    (1003, 32, 37),         # K below the owner count
    (2051, 96, 384),
    (6000, 256, 512),
    (9001, 192, 256),
]


@pytest.mark.parametrize("N,D,K", SHAPES)
def test_paths_agree_with_oracle(lib, N, D, K):
    E, z, g, idx = _make(N, D, K, seed=N + D + K)
    dz_ref, dE_ref = _oracle(g, 0.7, z, E, idx)
    outs = {}
    for name, fl in (("flat", FLAT), ("private", PRIVATE)):
        assert lib.vq_backward_path(N, K, D, fl) in (0, 2)
        dz, dE = _backward(lib, g, 0.7, z, E, idx, TRAIN | ZERO_DE | fl)
        outs[name] = (dz, dE, lib.vq_backward_path(N, K, D, fl))
        assert _rel(dz.cpu().numpy(), dz_ref) <= 1e-5, name
        assert _rel(dE.cpu().numpy(), dE_ref) <= 1e-5, name
    # the forced path really ran (all of SHAPES have D % 32 == 0 and a dE slice that fits shared memory)
    assert outs["private"][2] == 2
    assert torch.equal(outs["flat"][0], outs["private"][0])
    # the default choice is one of them and agrees as well
    dz, dE = _backward(lib, g, 0.7, z, E, idx, TRAIN | ZERO_DE)
    assert torch.equal(dz, outs["flat"][0]) and _rel(dE.cpu().numpy(), dE_ref) <= 1e-5


def test_default_path_choice(lib):
    assert lib.vq_backward_path(51456, 1024, 64, 0) == 0          # few rows per code: flat
    assert lib.vq_backward_path(1 << 20, 512, 64, 0) == 3         # sweep, few addresses: reds spread over copies of dE
    assert lib.vq_backward_path(1 << 20, 512, 64, PRIVATE) == 2   # forced: per-CTA copy of dE in shared memory
    assert lib.vq_backward_path(1 << 20, 1024, 64, 0) == 3
    assert lib.vq_backward_path(1 << 20, 1024, 64, FLAT) == 0
    assert lib.vq_backward_path(1 << 20, 8192, 256, 0) == 0       # many addresses: the flat scatter is HBM-bound
    assert lib.vq_backward_path(1000, 1000, 5, 0) == 0            # odd D: flat


@pytest.mark.parametrize("N,D,K", [(300000, 64, 512), (262144, 128, 512), (270001, 64, 1024)])
def test_step_pair_small_codebook_large_n(lib, N, D, K):
    """Large N with a small codebook: vq_step_forward keeps the usage counts per CTA in shared memory (and, under
    B200VQ_SPLIT_ROWS=1, runs quantize_rows_kernel behind an indices-only screen kernel), vq_step_backward spreads its reds
    over copies of dE in the workspace's dead tail.  Indices
    must equal the exact CUDA-core path's, dz the flat kernel's bit for bit, dE / loss / perplexity within tolerance; the
    workspace must be reusable by the next forward (two steps back to back)."""
    dev = _dev()
    st = torch.cuda.current_stream().cuda_stream
    assert lib.vq_backward_path(N, K, D, 0) == 3
    E = torch.randn(K, D, device=dev)
    e2 = torch.empty(K, device=dev); ehi = torch.empty_like(E); elo = torch.empty_like(E)
    wsb = lib.vq_workspace_bytes(N, K, D, 0); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    q = torch.empty(N, D, device=dev); idx = torch.empty(N, dtype=torch.int32, device=dev)
    stats = torch.empty(K + 3, device=dev); sp = stats.data_ptr()
    dz = torch.empty(N, D, device=dev); dE = torch.empty(K, D, device=dev)
    gl = torch.tensor(0.7, device=dev)
    for step in range(2):
        z = torch.randn(N, D, device=dev); g = torch.randn(N, D, device=dev)
        rc = lib.vq_step_forward(z.data_ptr(), E.data_ptr(), N, K, D, BETA, 0, e2.data_ptr(), ehi.data_ptr(), elo.data_ptr(), dE.data_ptr(),
                                 q.data_ptr(), idx.data_ptr(), None, sp, sp + 4 * K, sp + 4 * (K + 1), sp + 4 * (K + 2), ws.data_ptr(), wsb, st)
        assert rc == 0, lib.vq_last_error()
        rc = lib.vq_step_backward(g.data_ptr(), gl.data_ptr(), z.data_ptr(), E.data_ptr(), idx.data_ptr(), N, N, N, K, D, BETA, TRAIN,
                                  dz.data_ptr(), dE.data_ptr(), ws.data_ptr(), wsb, 0, st)
        assert rc == 0, lib.vq_last_error()
        torch.cuda.synchronize()
        # exact path (CUDA cores, oracle order) for the indices and the statistics
        q2 = torch.empty_like(q); idx2 = torch.empty_like(idx); stats2 = torch.empty_like(stats); sp2 = stats2.data_ptr()
        rc = lib.vq_step_forward(z.data_ptr(), E.data_ptr(), N, K, D, BETA, 1 << 2, e2.data_ptr(), ehi.data_ptr(), elo.data_ptr(), None,
                                 q2.data_ptr(), idx2.data_ptr(), None, sp2, sp2 + 4 * K, sp2 + 4 * (K + 1), sp2 + 4 * (K + 2), ws.data_ptr(), wsb, st)
        assert rc == 0, lib.vq_last_error()
        torch.cuda.synchronize()
        assert torch.equal(idx, idx2), f"step {step}: {int((idx != idx2).sum())} rows differ"
        assert torch.equal(q, q2)
        assert torch.equal(stats[:K], stats2[:K])                                   # usage counts
        assert abs(float(stats[K + 1]) - float(stats2[K + 1])) <= 1e-5 * float(stats2[K + 1])
        assert abs(float(stats[K + 2]) - float(stats2[K + 2])) <= 1e-5 * float(stats2[K + 2])
        dz_ref, dE_ref = _backward(lib, g, 0.7, z, E, idx, TRAIN | ZERO_DE | FLAT)
        assert torch.equal(dz, dz_ref)
        assert _rel(dE.cpu().numpy(), dE_ref.cpu().numpy()) <= 1e-5


@pytest.mark.parametrize("path", [FLAT, PRIVATE])
def test_hot_code_owns_every_row(lib, path):
    # one code takes all rows: every atomic hits the same row (flat) / one warp owns every row (private)
    N, D, K = 20000, 64, 512
    E, z, g, idx = _make(N, D, K, seed=5, hot=130)
    dz_ref, dE_ref = _oracle(g, 1.0, z, E, idx)
    dz, dE = _backward(lib, g, 1.0, z, E, idx, TRAIN | ZERO_DE | path)
    assert _rel(dz.cpu().numpy(), dz_ref) <= 1e-5
    assert _rel(dE.cpu().numpy(), dE_ref) <= 2e-5        # 20 000 addends in one fp32 sum
    assert float(dE[:130].abs().max()) == 0.0 and float(dE[131:].abs().max()) == 0.0


@pytest.mark.parametrize("path", [FLAT, PRIVATE])
def test_accumulate_and_overwrite(lib, path):
    N, D, K = 5000, 64, 256
    E, z, g, idx = _make(N, D, K, seed=9)
    _, dE_ref = _oracle(g, 1.0, z, E, idx)
    init = torch.randn(K, D, device=z.device)
    _, dE_over = _backward(lib, g, 1.0, z, E, idx, TRAIN | ZERO_DE | path, dE_init=init)       # dE = gradient
    _, dE_acc = _backward(lib, g, 1.0, z, E, idx, TRAIN | path, dE_init=init)                  # dE += gradient
    assert _rel(dE_over.cpu().numpy(), dE_ref) <= 1e-5
    # the flat path adds ~N/K small terms one by one into values of magnitude |init| <= ~4.5: a few ulp(4) = 4.8e-7 of rounding
    np.testing.assert_allclose((dE_acc - init).cpu().numpy(), dE_ref, rtol=0, atol=5e-6 + 1e-5 * np.abs(dE_ref).max())


@pytest.mark.parametrize("path", [FLAT, PRIVATE])
def test_codebook_gradient_only(lib, path):
    N, D, K = 7001, 64, 512
    E, z, g, idx = _make(N, D, K, seed=11)
    _, dE_ref = _oracle(g, 1.0, z, E, idx)
    dz, dE = _backward(lib, None, 1.0, z, E, idx, TRAIN | ZERO_DE | path, want_dz=False)
    assert dz is None and _rel(dE.cpu().numpy(), dE_ref) <= 1e-5


def test_global_row_count_scales_dE_only(lib):
    # data parallel: dE carries the GLOBAL row count in its scale, dz the local one
    N, D, K = 6432, 64, 1024
    E, z, g, idx = _make(N, D, K, seed=13)
    dz1, dE1 = _backward(lib, g, 1.0, z, E, idx, TRAIN | ZERO_DE)
    dz8, dE8 = _backward(lib, g, 1.0, z, E, idx, TRAIN | ZERO_DE, n_dE=8 * N)
    assert torch.equal(dz1, dz8)
    np.testing.assert_allclose(dE8.cpu().numpy() * 8, dE1.cpu().numpy(), rtol=0, atol=1e-5 * float(dE1.abs().max()))   # summation order differs between runs


def test_out_of_range_codes_are_ignored(lib):
    # indices outside [0, K) never come out of vq_forward; a caller-supplied one must not touch memory outside dE
    N, D, K = 4096, 64, 256
    E, z, g, idx = _make(N, D, K, seed=17)
    bad = idx.clone()
    bad[::97] = K + 5
    good_rows = (bad < K)
    for path in (PRIVATE,):
        guard = torch.zeros(K + 64, D, device=z.device)
        glt = torch.ones((), device=z.device)
        dz = torch.empty(N, D, device=z.device)
        st = torch.cuda.current_stream().cuda_stream
        rc = lib.vq_backward(g.data_ptr(), glt.data_ptr(), z.data_ptr(), E.data_ptr(), bad.data_ptr(), N, N, N, K, D, BETA,
                             TRAIN | ZERO_DE | path, dz.data_ptr(), guard.data_ptr(), st)
        assert rc == 0, lib.vq_last_error()
        torch.cuda.synchronize()
        assert float(guard[K:].abs().max()) == 0.0
        assert torch.equal(dz[~good_rows], g[~good_rows])       # no code, no commitment term
        ref = torch.zeros(K, D, device=z.device, dtype=torch.float64)
        ref.index_add_(0, bad[good_rows].long(), (E[bad[good_rows].long()] - z[good_rows]).double())
        ref *= 2.0 / (N * D)
        assert _rel(guard[:K].cpu().numpy(), ref.cpu().numpy()) <= 1e-5


def test_prepare_fast_matches_oracle_norms(lib):
    from oracle import c_oracle
    dev = _dev()
    for K, D in ((1024, 64), (1000, 128), (37, 32), (8192, 256), (129, 100)):
        E = torch.randn(K, D, device=dev)
        e2 = torch.empty(K, device=dev); ehi = torch.empty_like(E); elo = torch.empty_like(E)
        st = torch.cuda.current_stream().cuda_stream
        assert lib.vq_prepare_codebook(E.data_ptr(), K, D, e2.data_ptr(), ehi.data_ptr(), elo.data_ptr(), st) == 0
        torch.cuda.synchronize()
        assert np.array_equal(e2.cpu().numpy(), c_oracle.code_norms(E.cpu().numpy())), (K, D)
        # hi is E rounded to tf32 (10 explicit mantissa bits), lo the rounded remainder: hi + lo reproduces E to 2^-21
        hi = ehi.cpu().numpy(); lo = elo.cpu().numpy()
        assert np.all((hi.view(np.uint32) & 0x1FFF) == 0) and np.all((lo.view(np.uint32) & 0x1FFF) == 0)
        assert np.abs(hi + lo - E.cpu().numpy()).max() <= 2.0 ** -21 * np.abs(E.cpu().numpy()).max()


# ---- data-parallel exchange, `world` ranks emulated on one GPU ------------------------------------------------
@pytest.mark.parametrize("world,two_step", [(2, 0), (4, 0), (8, 1), (8, 0), (3, 1), (16, 1)])
@pytest.mark.parametrize("n", [1024 * 64 + 1024 + 1, 4097, 64])
def test_dp_exchange_emulated_ranks(lib, world, two_step, n):
    """vq_dp_emulate runs the production exchange code with blockIdx.y as the rank (one cooperative launch, all receive
    buffers on this GPU): every rank must end up with the rank-ordered sum, bit-identical across ranks, for several
    calls in a row (buffer parity, device-side sequence numbers)."""
    import ctypes
    dev = _dev()
    if two_step and 2 * ((((n + 1) // 2) + world - 1) // world) > (n + 1) // 2 + 2:
        pytest.skip("payload too small for two steps")
    g0 = torch.Generator(device="cpu").manual_seed(world * 1000 + n)
    pay = [torch.randn(n, generator=g0).to(dev) for _ in range(world)]
    outs = [torch.full((n,), -3.0, device=dev) for _ in range(world)]
    PA = (ctypes.c_void_p * world)(*[p.data_ptr() for p in pay])
    OA = (ctypes.c_void_p * world)(*[o.data_ptr() for o in outs])
    err = ctypes.c_uint32(0)
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.vq_dp_emulate(world, two_step, n, PA, OA, 5, 0, ctypes.byref(err), st)
    assert rc == 0, lib.vq_last_error()
    assert err.value == 0, f"error word {err.value:#x}"
    ref = torch.zeros(n, device=dev)
    for p in pay:                                           # rank order, like the kernel
        ref = ref + p
    for r in range(world):
        assert torch.equal(outs[r], ref), f"rank {r}: max diff {float((outs[r] - ref).abs().max())}"


def test_dp_exchange_bounded_wait(lib):
    """A rank whose peers never send must give up after spin_limit polls and raise the error word instead of hanging:
    emulated by a context of world 2 driven alone."""
    import ctypes
    dev = _dev()
    n = 4096
    lines = int(lib.vq_dp_recv_lines(2, n))
    bufs = [torch.zeros(lines * 4, device=dev) for _ in range(4)]       # [parity][rank], never written by "rank 1"
    R0 = (ctypes.c_void_p * 2)(bufs[0].data_ptr(), bufs[1].data_ptr())
    R1 = (ctypes.c_void_p * 2)(bufs[2].data_ptr(), bufs[3].data_ptr())
    ctx = ctypes.c_void_p()
    assert lib.vq_dp_create(R0, R1, None, None, 2, 0, n, 2000, ctypes.byref(ctx)) == 0, lib.vq_last_error()
    pay = torch.randn(n, device=dev); out = torch.zeros(n, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    assert lib.vq_dp_allreduce(ctx, pay.data_ptr(), out.data_ptr(), st) == 0, lib.vq_last_error()
    calls, err = ctypes.c_uint32(0), ctypes.c_uint32(0)
    assert lib.vq_dp_status(ctx, ctypes.byref(calls), ctypes.byref(err), st) == 0
    assert err.value & 1 == 1 and calls.value == 1
    lib.vq_dp_destroy(ctx)


def test_step_backward_behind_the_forward(lib):
    """vq_step_backward starts on the workspace's ready word, before the forward's statistics tail has finished: over many
    back-to-back steps with fresh inputs it must see complete indices every time (dz bit-identical to a backward run
    after a full synchronise, dE within tolerance), and the forward's loss / perplexity must be complete afterwards."""
    dev = _dev()
    for (N, D, K, onehot) in ((51456, 64, 1024, True), (16000, 128, 1024, False), (3216, 64, 1024, True), (300, 32, 256, False)):
        st = torch.cuda.current_stream().cuda_stream
        E = torch.randn(K, D, device=dev)
        e2 = torch.empty(K, device=dev); ehi = torch.empty_like(E); elo = torch.empty_like(E)
        wsb = lib.vq_workspace_bytes(N, K, D, 0); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        q = torch.empty(N, D, device=dev); idx = torch.empty(N, dtype=torch.int32, device=dev)
        oh = torch.empty(N, K, device=dev) if onehot else None
        stats = torch.empty(K + 3, device=dev); sp = stats.data_ptr()
        dz = torch.empty(N, D, device=dev); dE = torch.empty(K, D, device=dev)
        gl = torch.tensor(0.7, device=dev)
        fl = 1 if onehot else 0
        zs = [torch.randn(N, D, device=dev) for _ in range(6)]
        gs = [torch.randn(N, D, device=dev) for _ in range(6)]
        outs = []
        for i in range(6):                                   # enqueue everything, no synchronisation in between
            rc = lib.vq_step_forward(zs[i].data_ptr(), E.data_ptr(), N, K, D, BETA, fl, e2.data_ptr(), ehi.data_ptr(), elo.data_ptr(), dE.data_ptr(),
                                     q.data_ptr(), idx.data_ptr(), None if oh is None else oh.data_ptr(), sp, sp + 4 * K, sp + 4 * (K + 1),
                                     sp + 4 * (K + 2), ws.data_ptr(), wsb, st)
            assert rc == 0, lib.vq_last_error()
            rc = lib.vq_step_backward(gs[i].data_ptr(), gl.data_ptr(), zs[i].data_ptr(), E.data_ptr(), idx.data_ptr(), N, N, N, K, D, BETA, TRAIN,
                                      dz.data_ptr(), dE.data_ptr(), ws.data_ptr(), wsb, fl, st)
            assert rc == 0, lib.vq_last_error()
            outs.append((idx.clone(), dz.clone(), dE.clone(), stats.clone()))
        torch.cuda.synchronize()
        for i, (ix, dz_i, dE_i, st_i) in enumerate(outs):
            dz_ref, dE_ref = _backward(lib, gs[i], 0.7, zs[i], E, ix, TRAIN | ZERO_DE | FLAT)
            assert torch.equal(dz_i, dz_ref), f"N={N} step {i}"
            assert _rel(dE_i.cpu().numpy(), dE_ref.cpu().numpy()) <= 1e-5, f"N={N} step {i}"
            d = (zs[i] * zs[i]).sum(1, keepdim=True) + (E * E).sum(1) - 2 * zs[i] @ E.t()
            assert float((ix.long() != d.argmin(1)).float().mean()) < 1e-3          # same codes as the plain torch argmin (up to near-ties)
            m = float(((E[ix.long()] - zs[i]) ** 2).mean())
            assert abs(float(st_i[K + 1]) - 1.25 * m) <= 1e-5 * 1.25 * m and float(st_i[K + 2]) > 1.0


def test_step_backward_dp_world_of_one(lib):
    """vq_step_backward_dp = dE pass -> exchange -> dz pass on one stream.  With a world-of-one exchange context (this GPU's own
    receive buffers) the summed buffer must equal the packed buffer, dz the fused kernel's bit for bit, dE the flat kernel's
    within tolerance -- over several back-to-back steps (sequence numbers, buffer parity, the ready word) and for shapes
    that take the fallback (odd D)."""
    import ctypes
    dev = _dev()
    st = torch.cuda.current_stream().cuda_stream
    for (N, D, K, onehot) in ((51456, 64, 1024, True), (16000, 128, 1024, False), (1003, 30, 37, False)):
        n = K * D + K + 1
        lines = int(lib.vq_dp_recv_lines(1, n))
        bufs = [torch.zeros(lines * 4, device=dev) for _ in range(2)]
        arr = [(ctypes.c_void_p * 1)(b.data_ptr()) for b in bufs]
        ctx = ctypes.c_void_p()
        assert lib.vq_dp_create(arr[0], arr[1], None, None, 1, 0, n, 0, ctypes.byref(ctx)) == 0, lib.vq_last_error()
        E = torch.randn(K, D, device=dev)
        e2 = torch.empty(K, device=dev); ehi = torch.empty_like(E); elo = torch.empty_like(E)
        wsb = lib.vq_workspace_bytes(N, K, D, 0); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        q = torch.empty(N, D, device=dev); idx = torch.empty(N, dtype=torch.int32, device=dev)
        oh = torch.empty(N, K, device=dev) if onehot else None
        packed = torch.zeros(n + 2, device=dev); out = torch.zeros(n, device=dev)
        pk = packed.data_ptr(); sp = pk + 4 * K * D
        dz = torch.empty(N, D, device=dev)
        gl = torch.tensor(0.7, device=dev)
        fl = 1 if onehot else 0
        res = []
        zs = [torch.randn(N, D, device=dev) for _ in range(4)]
        gs = [torch.randn(N, D, device=dev) for _ in range(4)]
        for i in range(4):
            rc = lib.vq_step_forward(zs[i].data_ptr(), E.data_ptr(), N, K, D, BETA, fl, e2.data_ptr(), ehi.data_ptr(), elo.data_ptr(), pk,
                                     q.data_ptr(), idx.data_ptr(), None if oh is None else oh.data_ptr(), sp, sp + 4 * K, sp + 4 * (K + 1),
                                     sp + 4 * (K + 2), ws.data_ptr(), wsb, st)
            assert rc == 0, lib.vq_last_error()
            rc = lib.vq_step_backward_dp(gs[i].data_ptr(), gl.data_ptr(), zs[i].data_ptr(), E.data_ptr(), idx.data_ptr(), N, N, 2 * N, K, D, BETA, TRAIN,
                                         dz.data_ptr(), pk, ctx, out.data_ptr(), ws.data_ptr(), wsb, fl, st)
            assert rc == 0, lib.vq_last_error()
            res.append((idx.clone(), dz.clone(), packed[:n].clone(), out.clone()))
        torch.cuda.synchronize()
        calls, err = ctypes.c_uint32(0), ctypes.c_uint32(0)
        assert lib.vq_dp_status(ctx, ctypes.byref(calls), ctypes.byref(err), st) == 0
        assert calls.value == 4 and err.value == 0
        for i, (ix, dz_i, pk_i, out_i) in enumerate(res):
            assert torch.equal(pk_i, out_i), f"N={N} step {i}: world-of-one sum differs from the payload"
            dz_ref, dE_ref = _backward(lib, gs[i], 0.7, zs[i], E, ix, TRAIN | ZERO_DE | FLAT, n_dE=2 * N)
            assert torch.equal(dz_i, dz_ref), f"N={N} step {i}"
            assert _rel(out_i[:K * D].view(K, D).cpu().numpy(), dE_ref.cpu().numpy()) <= 1e-5, f"N={N} step {i}"
            hist = torch.bincount(ix.long(), minlength=K).float()
            assert torch.equal(out_i[K * D:K * D + K], hist)
        lib.vq_dp_destroy(ctx)
