"""GPU property tests at BASELINE.json's full sizes (configs[1] and the N = 1M sweep), where the CPU
oracle would take minutes: size-independent properties the domain offers.

  * the default tensor path (tcgen05 screen + refine, fused forward) agrees with the exact CUDA-core path (bit-exact
    vs oracle/vq_oracle.c at oracle-sized cases) on EVERY row; the 3xTF32 kernel agrees except on fp32 near-ties;
  * histogram / one-hot / index consistency (checksum of checksums);
  * idempotence: quantizing the codewords themselves returns their own indices and a zero loss;
  * the loss identity loss == (1+beta) * mean((q_out - z)^2) and the gradient sum rules
        sum_k dE_k == ce * sum_n (E[idx_n] - z_n),   dz - g == -beta * ce' * (E[idx] - z);
  * linearity of backward in (g_loss, g_q);
  * a slice of the big problem checked against the CPU oracle.
"""
import os

import numpy as np
import pytest
import torch

from util import NEAR_TIE_TENSOR

pytestmark = pytest.mark.gpu
BETA = 0.25


def _dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _forward(lib, z, E, flags, want_onehot=False, step=False):
    """step=False: vq_prepare_codebook + vq_forward.  step=True: vq_step_forward (prepare launch with the state / dE reset
    + forward behind one entry) -- called three times on the same workspace, every call must give the same outputs."""
    dev = z.device
    N, D = z.shape
    K = E.shape[0]
    st = torch.cuda.current_stream().cuda_stream
    e2 = torch.empty(K, device=dev); ehi = torch.empty_like(E); elo = torch.empty_like(E)
    q = torch.empty_like(z); idx = torch.empty(N, dtype=torch.int32, device=dev)
    oh = torch.empty(N, K, device=dev) if want_onehot else None
    stats = torch.empty(K + 3, device=dev)
    fl = flags | (1 if want_onehot else 0)
    wsb = lib.vq_workspace_bytes(N, K, D, fl); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    sp = stats.data_ptr()
    if not step:
        assert lib.vq_prepare_codebook(E.data_ptr(), K, D, e2.data_ptr(), ehi.data_ptr(), elo.data_ptr(), st) == 0
        rc = lib.vq_forward(z.data_ptr(), E.data_ptr(), e2.data_ptr(), ehi.data_ptr(), elo.data_ptr(), N, K, D, BETA, fl,
                            q.data_ptr(), idx.data_ptr(), None if oh is None else oh.data_ptr(), sp, sp + 4 * K, sp + 4 * (K + 1),
                            sp + 4 * (K + 2), ws.data_ptr(), wsb, st)
        assert rc == 0, lib.vq_last_error()
        torch.cuda.synchronize()
        return dict(q=q, idx=idx, onehot=oh, hist=stats[:K], sse=stats[K], loss=stats[K + 1], perplexity=stats[K + 2], e_norm2=e2)
    prev = None
    dEz = torch.full((K, D), 3.0, device=dev)
    for call in range(3):
        stats.fill_(-5.0); idx.fill_(-1); dEz.fill_(3.0)
        rc = lib.vq_step_forward(z.data_ptr(), E.data_ptr(), N, K, D, BETA, fl, e2.data_ptr(), ehi.data_ptr(), elo.data_ptr(), dEz.data_ptr(),
                                 q.data_ptr(), idx.data_ptr(), None if oh is None else oh.data_ptr(), sp, sp + 4 * K, sp + 4 * (K + 1),
                                 sp + 4 * (K + 2), ws.data_ptr(), wsb, st)
        assert rc == 0, lib.vq_last_error()
        torch.cuda.synchronize()
        cur = (idx.clone(), stats.clone())
        assert float(dEz.abs().max()) == 0.0                 # the prepare launch zeroed the accumulator
        if prev is not None:
            assert torch.equal(prev[0], cur[0]) and torch.equal(prev[1], cur[1]), f"vq_step_forward call {call} differs from call {call - 1}"
        prev = cur
    return dict(q=q, idx=idx, onehot=oh, hist=stats[:K], sse=stats[K], loss=stats[K + 1], perplexity=stats[K + 2], e_norm2=e2)


def _near_tie_gap(z, E, idx_a, idx_b):
    mis = torch.nonzero(idx_a != idx_b).flatten()
    if mis.numel() == 0:
        return 0, 0.0
    zr = z[mis].double()
    ea, eb = E[idx_a[mis].long()].double(), E[idx_b[mis].long()].double()
    da, db = ((zr - ea) ** 2).sum(1), ((zr - eb) ** 2).sum(1)
    scale = (zr ** 2).sum(1) + torch.maximum((ea ** 2).sum(1), (eb ** 2).sum(1))
    return int(mis.numel()), float(((da - db).abs() / scale).max())


FULL = [
    ("rir256", 256 * 201, 64, 1024, True),        # BASELINE configs[1], dense one-hot on
    ("sweep_k1024_d64", 1 << 20, 64, 1024, False),  # configs[3]
    ("sweep_k512_d64", 1 << 20, 64, 512, False),
    ("sweep_k2048_d128", 1 << 19, 128, 2048, False),
]


@pytest.mark.parametrize("name,N,D,K,onehot", FULL)
def test_full_size_properties(lib, name, N, D, K, onehot):
    from oracle import c_oracle
    dev = _dev()
    torch.manual_seed(0)
    E = torch.randn(K, D, device=dev)
    z = torch.randn(N, D, device=dev)
    t = _forward(lib, z, E, 0, want_onehot=onehot)           # tensor path, fused
    assert lib.vq_forward_uses_tensor_path(N, K, D, 0) == 1
    ts = _forward(lib, z, E, 0, want_onehot=onehot, step=True)   # the same through vq_step_forward
    assert torch.equal(ts["idx"], t["idx"]) and torch.equal(ts["hist"], t["hist"]) and torch.equal(ts["q"], t["q"])
    assert torch.equal(ts["e_norm2"], t["e_norm2"])
    assert abs(float(ts["loss"]) - float(t["loss"])) <= 1e-6 * float(t["loss"]) and abs(float(ts["perplexity"]) - float(t["perplexity"])) <= 1e-6 * float(t["perplexity"])
    x = _forward(lib, z, E, 4)                                # exact CUDA-core path
    n_mis, gap = _near_tie_gap(z, E, t["idx"], x["idx"])
    # these shapes run the screen + refine kernel, whose indices are bit-exact: not one row may differ at full size
    assert n_mis == 0, f"{name}: {n_mis} rows differ from the exact path, worst gap {gap:.2e}"
    y = _forward(lib, z, E, 1 << 9)                           # VQ_FLAG_NO_SCREEN: the 3xTF32 kernel, equal up to fp32 near-ties
    n_mis, gap = _near_tie_gap(z, E, y["idx"], x["idx"])
    assert gap <= NEAR_TIE_TENSOR and n_mis <= max(1, N // 2000), f"{name} (3xTF32): {n_mis} rows differ, worst gap {gap:.2e}"
    # histogram / indices / one-hot: a checksum of checksums
    hist = torch.bincount(t["idx"].long(), minlength=K).float()
    assert torch.equal(hist, t["hist"]) and float(t["hist"].sum()) == N
    assert int(t["idx"].min()) >= 0 and int(t["idx"].max()) < K
    if onehot:
        oh = t["onehot"]
        assert torch.equal(oh.sum(0), t["hist"]) and torch.equal(oh.argmax(1).int(), t["idx"])
        assert float(oh.sum()) == N and float(oh.max()) == 1.0 and float(oh.min()) == 0.0
    # straight-through value, loss and perplexity identities (fp64 on the GPU as the yardstick)
    qe = E[t["idx"].long()]
    assert torch.equal(t["q"], z + (qe - z))
    m = ((qe.double() - z.double()) ** 2).mean()
    assert abs(float(t["loss"]) - float(m * (1 + BETA))) <= 1e-5 * float(m * (1 + BETA))
    assert abs(float(t["sse"]) - float(m * N * D)) <= 1e-5 * float(m * N * D)
    p = (hist / N).double()
    perp = torch.exp(-(p * torch.log(p + 1e-10)).sum())
    assert abs(float(t["perplexity"]) - float(perp)) <= 1e-5 * float(perp)
    # a slice against the CPU oracle (bit-exact indices on the exact path)
    rows = z[:4096].cpu().numpy()
    o_idx = c_oracle.argmin(rows, E.cpu().numpy())
    assert np.array_equal(x["idx"][:4096].cpu().numpy(), o_idx)
    # idempotence: the codewords quantize to themselves with zero error
    s = _forward(lib, E.clone(), E, 0)
    dup = torch.equal(s["idx"], torch.arange(K, dtype=torch.int32, device=dev))
    assert dup and float(s["sse"]) == 0.0 and torch.equal(s["q"], E)


def test_backward_sum_rules_and_linearity(lib):
    dev = _dev()
    N, D, K = 256 * 201, 64, 1024
    torch.manual_seed(1)
    E = torch.randn(K, D, device=dev); z = torch.randn(N, D, device=dev); g = torch.randn(N, D, device=dev)
    idx = _forward(lib, z, E, 0)["idx"]
    st = torch.cuda.current_stream().cuda_stream

    def bwd(g_q, g_loss):
        dz = torch.empty_like(z); dE = torch.zeros_like(E)
        gl = torch.tensor(g_loss, device=dev)
        rc = lib.vq_backward(None if g_q is None else g_q.data_ptr(), gl.data_ptr(), z.data_ptr(), E.data_ptr(), idx.data_ptr(),
                             N, N, N, K, D, BETA, 2, dz.data_ptr(), dE.data_ptr(), st)
        assert rc == 0
        torch.cuda.synchronize()
        return dz, dE

    dz1, dE1 = bwd(g, 1.0)
    diff = (E[idx.long()] - z).double()
    ce = 2.0 / (N * D)
    assert torch.allclose(dE1.double().sum(0), ce * diff.sum(0), rtol=1e-4, atol=1e-9)
    ref_dE = torch.zeros(K, D, dtype=torch.float64, device=dev).index_add_(0, idx.long(), ce * diff)
    assert float((dE1.double() - ref_dE).abs().max()) <= 1e-5 * float(ref_dE.abs().max())
    assert float((dz1.double() - (g.double() - BETA * ce * diff)).abs().max()) <= 1e-6
    # linearity: backward(a*g, b*g_loss) == a*backward(g, 0) + b*backward(0, g_loss)
    dz_g, dE_g = bwd(g, 0.0)
    dz_l, dE_l = bwd(None, 1.0)
    dz2, dE2 = bwd(2.0 * g, 3.0)
    assert torch.allclose(dz2, 2.0 * dz_g + 3.0 * dz_l, rtol=1e-5, atol=1e-7)
    assert torch.allclose(dE2, 3.0 * dE_l, rtol=1e-5, atol=1e-9) and float(dE_g.abs().max()) == 0.0
    assert torch.equal(dz_g, g)        # straight-through: the upstream gradient passes unchanged


def test_uniform_init_ties_at_full_size(lib):
    """At the reference's own init (U(-1/K, 1/K), vector_quantizer.py:16) |z|^2 dwarfs |E|^2 and fp32 ties are
    common: the tensor path must still agree with the exact path up to near-ties."""
    dev = _dev()
    N, D, K = 256 * 201, 64, 1024
    torch.manual_seed(2)
    E = (torch.rand(K, D, device=dev) * 2 - 1) / K
    z = torch.randn(N, D, device=dev)
    t = _forward(lib, z, E, 0)
    x = _forward(lib, z, E, 4)
    assert torch.equal(t["idx"], x["idx"])                   # screen + refine resolves the ties exactly
    y = _forward(lib, z, E, 1 << 9)                          # 3xTF32: up to near-ties
    n_mis, gap = _near_tie_gap(z, E, y["idx"], x["idx"])
    assert gap <= NEAR_TIE_TENSOR and n_mis <= N // 200, (n_mis, gap)


def _oracle_idx(z, E):
    from oracle import c_oracle
    return torch.from_numpy(c_oracle.argmin(z.cpu().numpy(), E.cpu().numpy()))


ADVERSARIAL = ["duplicates", "clustered", "rows_are_codes", "huge_scale", "tiny_scale", "one_dominant_code", "few_rows", "constant_rows",
               "six_copies", "hidden_pairs"]


@pytest.mark.parametrize("kind", ADVERSARIAL)
@pytest.mark.parametrize("D,K", [(64, 512), (128, 1024), (256, 256)])
def test_screen_kernel_adversarial_inputs(lib, kind, D, K):
    """The screen + refine kernel must stay BIT-EXACT vs oracle/vq_oracle.c when its rare paths fire: exact ties
    (duplicated codewords: first index wins), clustered codebooks (many candidates per row: chain-instance
    rescans, pair-list overflow, whole-codebook rescans), rows equal to codewords, extreme scales."""
    dev = _dev()
    g = torch.Generator().manual_seed(hash((kind, D, K)) % (1 << 31))
    N = 1500 if kind != "few_rows" else 3
    E = torch.randn(K, D, generator=g)
    z = torch.randn(N, D, generator=g)
    if kind == "duplicates":
        E[5::7] = E[3:3 + len(E[5::7])]                 # many exact duplicates: ties must go to the lower index
        z[:200] = E[torch.randint(0, K, (200,), generator=g)] + 1e-3 * torch.randn(200, D, generator=g)
    elif kind == "clustered":
        base = torch.randn(8, D, generator=g)
        E = base[torch.randint(0, 8, (K,), generator=g)] + 2e-4 * torch.randn(K, D, generator=g)   # ~K/8 near-ties per row
    elif kind == "rows_are_codes":
        z = E[torch.randint(0, K, (N,), generator=g)].clone()
    elif kind == "huge_scale":
        E, z = E * 3e3, z * 3e3
    elif kind == "tiny_scale":
        E, z = E * 1e-4, z * 1e-4
    elif kind == "one_dominant_code":
        E[17] *= 50.0                                   # inflates max|E_k| and with it every row's margin
    elif kind == "constant_rows":
        z[:] = z[0]
    elif kind == "six_copies":
        # six identical codewords spread over tiles and column halves: 6 exact ties per nearby row, more than the 4-entry
        # candidate log and handoff hold -> the global spill list; the lowest index must win
        copies = [(3 + j * (K // 6 + 1)) % K for j in range(6)]
        E[copies] = E[copies[0]].clone()
        z[:300] = E[copies[0]] + 1e-3 * torch.randn(300, D, generator=g)
    elif kind == "hidden_pairs":
        # two chain-instances (same tile half, columns equal mod 4) that each hide a second tied codeword behind their
        # best: the second place to look goes through the spill list
        E[7] = E[3].clone()
        hi = K // 2 + 88
        E[hi] = E[3].clone(); E[hi + 4] = E[3].clone()
        z[:300] = E[3] + 1e-3 * torch.randn(300, D, generator=g)
    E, z = E.contiguous().to(dev), z.contiguous().to(dev)
    out = _forward(lib, z, E, 1 << 10)                  # VQ_FLAG_SCREEN
    ref = _oracle_idx(z, E)
    assert torch.equal(out["idx"].cpu(), ref), f"{kind} D={D} K={K}: {(out['idx'].cpu() != ref).sum().item()} rows differ from the oracle"
    assert float(out["hist"].sum()) == N
    # the same through the one-call step entry
    out_s = _forward(lib, z, E, 1 << 10, step=True)
    assert torch.equal(out_s["idx"].cpu(), ref), f"{kind} D={D} K={K} (vq_step_forward): {(out_s['idx'].cpu() != ref).sum().item()} rows differ"
    assert torch.equal(out_s["hist"], out["hist"])
    out_oh = _forward(lib, z, E, 1 << 10, want_onehot=True)
    assert torch.equal(out_oh["idx"].cpu(), ref) and torch.equal(out_oh["onehot"].argmax(1).int().cpu(), ref)
    assert float(out_oh["onehot"].sum()) == N


def test_screen_kernel_fuzz_against_exact_path(lib):
    """Randomised sweep over shapes, scales and codebook geometries: the screen + refine kernel (append-only candidate
    log, compaction, spill list, rescans) must return exactly the indices of the exact CUDA-core path (which is
    bit-exact vs oracle/vq_oracle.c) -- with and without the dense one-hot (one / two row-worker groups)."""
    dev = _dev()
    rng = np.random.default_rng(20251018)
    n_checked = 0
    n_trials = int(os.environ.get("B200VQ_FUZZ_TRIALS", "48"))      # raise for a long soak run
    for trial in range(n_trials):
        D = int(rng.choice([32, 64, 96, 128, 192, 256]))
        K = int(rng.choice([256, 512, 768, 1024, 2048]))
        N = int(rng.integers(1, 6000))
        if trial % 4 == 3:                               # several 256-row items per CTA pair: slot / spill-list rotation,
            N = int(rng.integers(40000, 200000))         # alternating worker groups, ragged last item
            K = min(K, 1024)
        g = torch.Generator().manual_seed(int(rng.integers(1 << 30)))
        kind = trial % 6
        E = torch.randn(K, D, generator=g)
        z = torch.randn(N, D, generator=g)
        if kind == 1:                                    # reference init: tiny codebook, ties galore
            E = (torch.rand(K, D, generator=g) * 2 - 1) / K
        elif kind == 2:                                  # clusters of near-identical codewords, noise level sets the
            n_cl = int(rng.choice([4, 16, 64]))          # number of candidates per row (1 ... hundreds)
            noise = float(rng.choice([1e-5, 1e-4, 1e-3, 1e-2]))
            base = torch.randn(n_cl, D, generator=g)
            E = base[torch.randint(0, n_cl, (K,), generator=g)] + noise * torch.randn(K, D, generator=g)
        elif kind == 3:                                  # rows sit next to codewords, several exact copies of some codes
            E[torch.randint(0, K, (K // 8,), generator=g)] = E[torch.randint(0, K, (K // 8,), generator=g)]
            z = E[torch.randint(0, K, (N,), generator=g)] + 1e-3 * torch.randn(N, D, generator=g)
        elif kind == 4:                                  # wild scales
            s = float(rng.choice([1e-3, 30.0, 1e3]))
            E, z = E * s, z * s
        elif kind == 5:                                  # a few dominant codewords inflate the margin of every row
            E[torch.randint(0, K, (3,), generator=g)] *= float(rng.choice([5.0, 40.0]))
        E, z = E.contiguous().to(dev), z.contiguous().to(dev)
        exact = _forward(lib, z, E, 4)["idx"]
        for want_onehot in (False, True):
            if want_onehot and N * K * 4 > (1 << 30):
                continue
            for step in (False, True):
                out = _forward(lib, z, E, 1 << 10, want_onehot=want_onehot, step=step)
                bad = int((out["idx"] != exact).sum())
                assert bad == 0, f"trial {trial}: kind {kind} N={N} K={K} D={D} onehot={want_onehot} step={step}: {bad} rows differ"
                assert float(out["hist"].sum()) == N
            n_checked += 1
    assert n_checked >= 48
