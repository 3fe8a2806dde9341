"""GPU parity tests (run with `-m gpu` on a B200): the CUDA path, called through the C ABI and
through the nn.Module mirror, against the CPU oracle and the golden fixtures generated from the
real reference class.

Bars (BASELINE.json north_star): code indices identical except documented fp32 near-ties;
quantized / loss / perplexity (and dz / dE) within 1e-5 relative.
  * exact path  (VQ_FLAG_EXACT, CUDA cores, oracle FMA-chain order): indices BIT-EXACT vs
    oracle/vq_oracle.c on every case.
  * tensor path, default (tcgen05 TF32 screen + exact fp32 refine): indices BIT-EXACT vs the oracle as well;
    the 3xTF32 kernels (other shapes, VQ_FLAG_NO_SCREEN): every mismatching row must be an fp32 near-tie (util.py).
"""
import numpy as np
import pytest
import torch

from cases import CASES, make_inputs
from util import NEAR_TIE_EXACT, NEAR_TIE_TENSOR, near_tie_report, rel_err

pytestmark = pytest.mark.gpu

RTOL = 1e-5   # north_star tolerance for floating-point outputs


def _dev():
    assert torch.cuda.is_available(), "GPU tests need a B200"
    return torch.device("cuda:0")


def run_abi(lib, z_rows, E, g_rows, g_loss, beta, train_vq, flags_extra=0, want_onehot=True):
    """Drive prepare -> forward -> backward through the raw C ABI on device buffers."""
    dev = _dev()
    N, D = z_rows.shape
    K = E.shape[0]
    z = torch.from_numpy(np.ascontiguousarray(z_rows)).to(dev)
    Ed = torch.from_numpy(np.ascontiguousarray(E)).to(dev)
    g = None if g_rows is None else torch.from_numpy(np.ascontiguousarray(g_rows)).to(dev)
    e2 = torch.empty(K, device=dev)
    ehi = torch.empty(K, D, device=dev)
    elo = torch.empty(K, D, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    assert lib.vq_prepare_codebook(Ed.data_ptr(), K, D, e2.data_ptr(), ehi.data_ptr(), elo.data_ptr(), st) == 0, lib.vq_last_error()
    q = torch.empty(N, D, device=dev)
    idx = torch.full((N,), -1, dtype=torch.int32, device=dev)
    onehot = torch.full((N, K), 7.0, device=dev) if want_onehot else None
    hist = torch.full((K,), 9.0, device=dev)
    scal = torch.full((3,), -1.0, device=dev)
    flags = flags_extra | (1 if want_onehot else 0)
    wsb = lib.vq_workspace_bytes(N, K, D, flags)
    ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=dev)
    rc = lib.vq_forward(z.data_ptr(), Ed.data_ptr(), e2.data_ptr(), ehi.data_ptr(), elo.data_ptr(), N, K, D, beta, flags,
                        q.data_ptr(), idx.data_ptr(), None if onehot is None else onehot.data_ptr(), hist.data_ptr(),
                        scal.data_ptr(), scal.data_ptr() + 4, scal.data_ptr() + 8, ws.data_ptr(), wsb, st)
    assert rc == 0, lib.vq_last_error()
    dz = torch.empty(N, D, device=dev)
    dE = torch.zeros(K, D, device=dev) if train_vq else None
    gl = torch.tensor(g_loss, dtype=torch.float32, device=dev)
    rc = lib.vq_backward(None if g is None else g.data_ptr(), gl.data_ptr(), z.data_ptr(), Ed.data_ptr(), idx.data_ptr(),
                         N, N, N, K, D, beta, 2 if train_vq else 0, dz.data_ptr(), None if dE is None else dE.data_ptr(), st)
    assert rc == 0, lib.vq_last_error()
    torch.cuda.synchronize()
    return dict(idx=idx.cpu().numpy(), q=q.cpu().numpy(), onehot=onehot, hist=hist.cpu().numpy(), sse=float(scal[0]),
                loss=float(scal[1]), perplexity=float(scal[2]), dz=dz.cpu().numpy(),
                dE=None if dE is None else dE.cpu().numpy(), e_norm2=e2.cpu().numpy())


def check_against_oracle_and_golden(out, name, golden, exact):
    from oracle import c_oracle
    arrays, meta = golden
    c = CASES[name]
    m = meta["cases"][name]
    E, z, g = make_inputs(c)
    D, K = c["D"], c["K"]
    rows = z.numpy().reshape(-1, D)
    En = E.numpy()
    N = rows.shape[0]
    # --- indices -------------------------------------------------------------------------------------
    o_idx = c_oracle.argmin(rows, En)
    ref_idx = arrays[f"{name}/idx"].astype(np.int32)
    if exact:
        assert np.array_equal(out["idx"], o_idx), f"{name}: exact path differs from oracle/vq_oracle.c on {(out['idx'] != o_idx).sum()} rows"
        assert np.array_equal(out["e_norm2"], c_oracle.code_norms(En)), "|E_k|^2 must be bit-exact"
    tol = NEAR_TIE_EXACT if exact else NEAR_TIE_TENSOR
    n_mis, gap = near_tie_report(rows, En, out["idx"], ref_idx)
    assert gap <= tol, f"{name}: {n_mis} rows differ from the reference, worst fp64 gap {gap:.3e} > {tol:.3e}"
    assert n_mis <= max(1, N // 500), f"{name}: {n_mis} near-tie rows out of {N}"
    # --- everything downstream of the indices, checked against the oracle run on OUR indices --------
    fwd = c_oracle.quantize(rows, En, out["idx"], c["beta"])
    assert rel_err(out["q"], fwd["quantized"]) <= 1e-6
    assert np.array_equal(out["hist"], fwd["hist"]) and out["hist"].sum() == N
    assert abs(out["loss"] - fwd["loss"]) <= RTOL * abs(fwd["loss"])
    assert abs(out["perplexity"] - fwd["perplexity"]) <= RTOL * abs(fwd["perplexity"])
    assert abs(out["sse"] - fwd["sse"]) <= RTOL * abs(fwd["sse"])
    if out["onehot"] is not None:
        oh = out["onehot"]
        assert torch.equal(oh.argmax(1).int().cpu(), torch.from_numpy(out["idx"]))
        assert float(oh.sum()) == N and float(oh.max()) == 1.0 and float(oh.min()) == (0.0 if K > 1 else 1.0)
    dz, dE = c_oracle.backward(g.numpy().reshape(-1, D), c["g_loss"], rows, En, out["idx"], c["beta"], c["train_vq"])
    assert rel_err(out["dz"], dz) <= RTOL
    if c["train_vq"]:
        assert rel_err(out["dE"], dE) <= RTOL
    else:
        assert out["dE"] is None
    # --- and against the reference's own numbers -------------------------------------------------------
    if n_mis == 0:
        assert abs(out["loss"] - m["loss"]) <= RTOL * abs(m["loss"])
        assert abs(out["perplexity"] - m["perplexity"]) <= RTOL * abs(m["perplexity"])
        if c["store_full"]:
            assert rel_err(out["q"], arrays[f"{name}/q"].reshape(-1, D)) <= 1e-6
            assert rel_err(out["dz"], arrays[f"{name}/dz"].reshape(-1, D)) <= RTOL
            if c["train_vq"]:
                assert rel_err(out["dE"], arrays[f"{name}/dE"]) <= RTOL
        else:
            r = arrays[f"{name}/rows"]
            assert rel_err(out["q"][r], arrays[f"{name}/q_rows"]) <= 1e-6
            assert rel_err(out["dz"][r], arrays[f"{name}/dz_rows"]) <= RTOL
            if c["train_vq"]:
                assert rel_err(out["dE"].sum(1), arrays[f"{name}/dE_rowsum"]) <= 2e-5
    else:
        # a near-tie row changes the loss by at most its (tiny) distance gap
        assert abs(out["loss"] - m["loss"]) <= 1e-4 * abs(m["loss"])
    return n_mis, gap


@pytest.mark.parametrize("name", list(CASES))
def test_exact_path_abi(lib, golden, name):
    c = CASES[name]
    E, z, g = make_inputs(c)
    D = c["D"]
    out = run_abi(lib, z.numpy().reshape(-1, D), E.numpy(), g.numpy().reshape(-1, D), c["g_loss"], c["beta"],
                  c["train_vq"], flags_extra=4)
    check_against_oracle_and_golden(out, name, golden, exact=True)


TENSOR_CASES = [n for n, c in CASES.items() if c["D"] in (32, 64, 96, 128) and c["K"] % 128 == 0]
SCREEN_CASES = [n for n, c in CASES.items() if c["D"] in (32, 64, 96, 128, 192, 256) and c["K"] % 256 == 0]
FLAG_TC_1CTA, FLAG_NO_FUSE, FLAG_NO_SCREEN, FLAG_SCREEN = 1 << 6, 1 << 7, 1 << 9, 1 << 10


def _run_case(lib, name, flags):
    c = CASES[name]
    E, z, g = make_inputs(c)
    D = c["D"]
    return run_abi(lib, z.numpy().reshape(-1, D), E.numpy(), g.numpy().reshape(-1, D), c["g_loss"], c["beta"],
                   c["train_vq"], flags_extra=flags)


@pytest.mark.parametrize("name", SCREEN_CASES)
def test_screen_refine_path_is_bit_exact(lib, golden, name):
    """Default tensor path: one TF32 screening pass + exact fp32 refine of the candidates (vq_screen_kernel).
    Indices must equal oracle/vq_oracle.c bit for bit -- not merely up to near-ties -- for D up to 256."""
    c = CASES[name]
    N = int(np.prod(c["shape"])) // c["D"]
    assert lib.vq_forward_uses_tensor_path(N, c["K"], c["D"], 0) == 1
    out = _run_case(lib, name, FLAG_SCREEN)                        # with the dense one-hot (run_abi asks for it)
    check_against_oracle_and_golden(out, name, golden, exact=True)
    E, z, g = make_inputs(c)
    out = run_abi(lib, z.numpy().reshape(-1, c["D"]), E.numpy(), g.numpy().reshape(-1, c["D"]), c["g_loss"], c["beta"],
                  c["train_vq"], flags_extra=0, want_onehot=False)     # the default choice without a one-hot
    check_against_oracle_and_golden(out, name, golden, exact=True)


@pytest.mark.parametrize("name", [n for n in TENSOR_CASES if CASES[n]["K"] % 256 == 0])
def test_3xtf32_fused_kernel(lib, golden, name):
    """The 3xTF32 persistent CTA-pair kernel with the fused row epilogue (VQ_FLAG_NO_SCREEN)."""
    check_against_oracle_and_golden(_run_case(lib, name, FLAG_NO_SCREEN), name, golden, exact=False)


@pytest.mark.parametrize("name", [n for n in TENSOR_CASES if CASES[n]["K"] % 256 == 0])
def test_3xtf32_unfused_kernels(lib, golden, name):
    """CTA-pair argmin kernel + separate rows kernel (VQ_FLAG_NO_FUSE): the path taken when the codebook is split."""
    check_against_oracle_and_golden(_run_case(lib, name, FLAG_NO_FUSE), name, golden, exact=False)


@pytest.mark.parametrize("name", [n for n in TENSOR_CASES if CASES[n]["K"] % 256 == 0])
def test_3xtf32_single_cta_kernel(lib, golden, name):
    """The M=128/N=128 single-CTA tcgen05 kernel (used when K % 256 != 0), forced with VQ_FLAG_TC_1CTA."""
    check_against_oracle_and_golden(_run_case(lib, name, FLAG_TC_1CTA), name, golden, exact=False)


@pytest.mark.parametrize("name", TENSOR_CASES)
def test_tensor_path_abi(lib, golden, name):
    """Whatever vq_forward picks by default for tensor-eligible shapes."""
    c = CASES[name]
    N = int(np.prod(c["shape"])) // c["D"]
    assert lib.vq_forward_uses_tensor_path(N, c["K"], c["D"], 0) == 1
    n_mis, gap = check_against_oracle_and_golden(_run_case(lib, name, 0), name, golden, exact=False)
    print(f"[tensor path] {name}: {n_mis} near-tie rows, worst relative fp64 gap {gap:.2e}")


@pytest.mark.parametrize("name", ["rir32_normal", "small_frozen", "odd_shape", "speech_uniform", "pooled"])
@pytest.mark.parametrize("exact", [True, False])
def test_module_autograd(golden, name, exact):
    """The nn.Module mirror under autograd: same call sequence as convolutional_vq_vae.py:98 +
    `(g_loss*loss + (g*q).sum()).backward()`; in-place edits of `quantized` (jitter.py:68) are legal."""
    import b200vq
    from oracle import c_oracle
    arrays, meta = golden
    c = CASES[name]
    m = meta["cases"][name]
    dev = _dev()
    E, z, g = make_inputs(c)
    vq = b200vq.VectorQuantizer(c["K"], c["D"], c["beta"], exact=exact).to(dev)
    vq._embedding.weight.data.copy_(E)
    vq.set_train_vq(c["train_vq"])
    zd = z.to(dev).requires_grad_(True)
    loss, q, perp, enc = vq(zd)
    assert loss.shape == () and perp.shape == () and q.shape == zd.shape and q.is_contiguous()
    assert loss.requires_grad and q.requires_grad and not perp.requires_grad and not enc.requires_grad
    N = enc.shape[0]
    assert enc.shape == (N, c["K"]) and enc.dtype == torch.float32
    obj = c["g_loss"] * loss + (g.to(dev) * q).sum()
    q_val = q.detach().clone()
    with torch.no_grad():
        q.mul_(0.0)     # Jitter mutates the returned tensor in place; backward must not depend on it
    obj.backward()
    torch.cuda.synchronize()
    idx = vq.last_indices.cpu().numpy()
    ref_idx = arrays[f"{name}/idx"].astype(np.int32)
    rows = z.numpy().reshape(-1, c["D"])
    n_mis, gap = near_tie_report(rows, E.numpy(), idx, ref_idx)
    assert gap <= (NEAR_TIE_EXACT if exact else NEAR_TIE_TENSOR)
    assert torch.equal(enc.argmax(1).int().cpu(), torch.from_numpy(idx))
    dz, dE = c_oracle.backward(g.numpy().reshape(-1, c["D"]), c["g_loss"], rows, E.numpy(), idx, c["beta"], c["train_vq"])
    assert rel_err(zd.grad.cpu().numpy().reshape(-1, c["D"]), dz) <= RTOL
    if c["train_vq"]:
        assert rel_err(vq._embedding.weight.grad.cpu().numpy(), dE) <= RTOL
    else:
        assert vq._embedding.weight.grad is None
    if n_mis == 0:
        assert abs(float(loss) - m["loss"]) <= RTOL * abs(m["loss"])
        assert abs(float(perp) - m["perplexity"]) <= RTOL * abs(m["perplexity"])
        if c["store_full"]:
            assert rel_err(q_val.cpu().numpy(), arrays[f"{name}/q"]) <= 1e-6


def test_module_options_and_errors():
    import b200vq
    dev = _dev()
    vq = b200vq.VectorQuantizer(128, 64, 0.25, return_encodings=False).to(dev)
    z = torch.randn(4, 64, 10, device=dev)
    loss, q, perp, enc = vq(z)
    assert enc is None and vq.last_indices.shape == (40,)
    oh = vq.encodings_from_indices()
    assert oh.shape == (40, 128) and torch.equal(oh.argmax(1).int(), vq.last_indices) and float(oh.sum()) == 40
    assert float(vq.usage_histogram().sum()) == 40
    # reference error behaviour (vector_quantizer.py:32)
    with pytest.raises(RuntimeError):
        vq(torch.randn(4, 10, 64, device=dev).permute(0, 2, 1))
    with pytest.raises(RuntimeError):
        vq(torch.randn(3, 5, 7, device=dev))
    with pytest.raises(RuntimeError, match="float32"):
        vq(z.double())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vq(z.cpu())
    # straight-through only: upstream gradient passes unchanged, nothing reaches a frozen codebook
    vq.set_train_vq(False)
    z2 = z.clone().requires_grad_(True)
    _, q2, _, _ = vq(z2)
    gq = torch.randn_like(q2)
    q2.backward(gq)
    assert torch.equal(z2.grad, gq) and vq._embedding.weight.grad is None
    # empty batch: the reference's `.view(-1, D)` cannot infer -1 for zero elements and raises
    with pytest.raises(RuntimeError):
        b200vq.VectorQuantizer(128, 64, 0.25).to(dev)(torch.empty(0, 64, 5, device=dev))


def test_host_buffer_entry_points(lib, golden):
    """vq_host_* (host pointers in, host pointers out) == device-pointer path == oracle."""
    import ctypes
    from oracle import c_oracle
    _dev()
    c = CASES["rir32_normal"]
    E, z, g = make_inputs(c)
    D, K = c["D"], c["K"]
    rows = np.ascontiguousarray(z.numpy().reshape(-1, D))
    N = rows.shape[0]
    ctx = ctypes.c_void_p()
    assert lib.vq_host_ctx_create(N, K, D, ctypes.byref(ctx)) == 0, lib.vq_last_error()
    try:
        En = np.ascontiguousarray(E.numpy())
        assert lib.vq_host_set_codebook(ctx, En.ctypes.data) == 0, lib.vq_last_error()
        loss = np.zeros(1, np.float32); perp = np.zeros(1, np.float32)
        idx = np.zeros(N, np.int32); q = np.zeros((N, D), np.float32); dz = np.zeros((N, D), np.float32)
        dE = np.zeros((K, D), np.float32)
        for lane, gq in ((0, None), (1, np.ascontiguousarray(g.numpy().reshape(-1, D)))):
            rc = lib.vq_host_step_async(ctx, lane, rows.ctypes.data, None if gq is None else gq.ctypes.data, N, 0, 0.25,
                                        2 | 4, loss.ctypes.data, perp.ctypes.data, idx.ctypes.data, q.ctypes.data,
                                        dz.ctypes.data, dE.ctypes.data)
            assert rc == 0, lib.vq_last_error()
            assert lib.vq_host_wait(ctx, lane) == 0
            ref = c_oracle.step(rows, En, np.ones_like(rows) if gq is None else gq, 1.0, 0.25, True)
            assert np.array_equal(idx, ref["indices"])
            assert abs(loss[0] - ref["loss"]) <= RTOL * ref["loss"] and abs(perp[0] - ref["perplexity"]) <= RTOL * ref["perplexity"]
            assert rel_err(q, ref["quantized"]) <= 1e-6 and rel_err(dz, ref["dz"]) <= RTOL and rel_err(dE, ref["dE"]) <= RTOL
    finally:
        lib.vq_host_ctx_destroy(ctx)
