"""CPU tests: the oracles against the golden fixtures generated from the real reference class
(tests/golden/make_golden.py), against SURVEY.md's known-answer test, and -- when /root/reference
is present -- against the reference itself."""
import numpy as np
import pytest
import torch

from cases import CASES, make_inputs
from oracle import c_oracle, vq_oracle
from util import NEAR_TIE_EXACT, near_tie_report, rel_err, sha16

SMALL = [n for n, c in CASES.items() if np.prod(c["shape"]) // c["D"] <= 7000]


def test_golden_inputs_reproducible(golden):
    """cases.make_inputs regenerates exactly the tensors the fixtures were made from."""
    arrays, meta = golden
    for name, c in CASES.items():
        E, z, g = make_inputs(c)
        m = meta["cases"][name]
        assert sha16(z.numpy()) == m["z_sha"], name
        assert sha16(E.numpy()) == m["E_sha"], name
        if c["store_full"]:
            assert np.array_equal(arrays[f"{name}/z"], z.numpy())
            assert np.array_equal(arrays[f"{name}/E"], E.numpy())


def test_survey_kat_tiny():
    """SURVEY.md Appendix A.1, produced from the reference: idx, loss, perplexity, gradients."""
    torch.manual_seed(1234)
    E = torch.empty(8, 4).normal_()
    E.uniform_(-1 / 8, 1 / 8)
    z = torch.randn(2, 4, 3)
    assert abs(float(E[0, 0]) - 0.07232261) < 1e-7 and abs(float(z.flatten()[0]) - 0.92574239) < 1e-6
    w = torch.arange(24.0).view(2, 4, 3)
    res = vq_oracle.forward_backward_dense(z, E, 0.25, g_quantized=w, g_loss=1.0)
    assert res.indices.tolist() == [4, 2, 3, 3, 2, 2]
    assert abs(float(res.loss) - 0.6078158020973206) < 1e-6
    assert abs(float(res.perplexity) - 2.7494592666625977) < 1e-5
    np.testing.assert_allclose(res.dE.sum(1).numpy(), [0, 0, 0.18623249, 0.34427917, -0.13527755, 0, 0, 0], atol=1e-6)
    np.testing.assert_allclose(res.dz[0, :, 0].numpy(), [0.017286995, 2.9963875, 6.0041599, 8.9930239], rtol=1e-5)
    # the C restatement agrees
    rows = z.numpy().reshape(-1, 4)
    idx = c_oracle.argmin(rows, E.numpy())
    assert idx.tolist() == [4, 2, 3, 3, 2, 2]
    qz = c_oracle.quantize(rows, E.numpy(), idx, 0.25)
    assert abs(qz["loss"] - 0.6078158020973206) < 1e-6 and abs(qz["perplexity"] - 2.7494592666625977) < 1e-5
    dz, dE = c_oracle.backward(w.numpy().reshape(-1, 4), 1.0, rows, E.numpy(), idx, 0.25)
    np.testing.assert_allclose(dE.sum(1), [0, 0, 0.18623249, 0.34427917, -0.13527755, 0, 0, 0], atol=1e-6)
    np.testing.assert_allclose(dz.reshape(2, 4, 3)[0, :, 0], [0.017286995, 2.9963875, 6.0041599, 8.9930239], rtol=1e-5)


@pytest.mark.parametrize("name", list(CASES))
def test_c_oracle_vs_golden(golden, name):
    """oracle/vq_oracle.c vs the reference's outputs: indices identical except fp32 near-ties;
    loss / perplexity / quantized / dz / dE within 1e-5 relative."""
    arrays, meta = golden
    c = CASES[name]
    m = meta["cases"][name]
    E, z, g = make_inputs(c)
    D = c["D"]
    rows = z.numpy().reshape(-1, D)
    En = E.numpy()
    idx = c_oracle.argmin(rows, En)
    ref_idx = arrays[f"{name}/idx"].astype(np.int32)
    n_mis, gap = near_tie_report(rows, En, idx, ref_idx)
    assert gap <= NEAR_TIE_EXACT, f"{name}: {n_mis} mismatching rows, worst fp64 gap {gap:.3e}"
    assert n_mis <= max(1, len(idx) // 500), f"{name}: {n_mis} near-tie rows"
    fwd = c_oracle.quantize(rows, En, ref_idx, c["beta"], want_onehot=len(idx) * c["K"] < 1 << 22)
    assert abs(fwd["loss"] - m["loss"]) <= 1e-5 * abs(m["loss"])
    assert abs(fwd["perplexity"] - m["perplexity"]) <= 1e-5 * abs(m["perplexity"])
    if fwd["onehot"] is not None:
        assert np.array_equal(fwd["onehot"].argmax(1), ref_idx) and fwd["onehot"].sum() == len(idx)
    dz, dE = c_oracle.backward(g.numpy().reshape(-1, D), c["g_loss"], rows, En, ref_idx, c["beta"], c["train_vq"])
    assert (dE is None) == m["dE_is_none"]
    if c["store_full"]:
        assert rel_err(fwd["quantized"], arrays[f"{name}/q"].reshape(-1, D)) <= 1e-6
        assert rel_err(dz, arrays[f"{name}/dz"].reshape(-1, D)) <= 1e-5
        if dE is not None:
            assert rel_err(dE, arrays[f"{name}/dE"]) <= 1e-5
    else:
        r = arrays[f"{name}/rows"]
        assert rel_err(fwd["quantized"][r], arrays[f"{name}/q_rows"]) <= 1e-6
        assert rel_err(dz[r], arrays[f"{name}/dz_rows"]) <= 1e-5
        if dE is not None:
            assert rel_err(dE.sum(1), arrays[f"{name}/dE_rowsum"]) <= 2e-5
            assert rel_err(dE.sum(0), arrays[f"{name}/dE_colsum"]) <= 2e-5


@pytest.mark.parametrize("name", SMALL)
def test_torch_oracle_vs_golden(golden, name):
    """oracle/vq_oracle.py (same aten ops as the reference) vs the fixtures."""
    arrays, meta = golden
    c = CASES[name]
    m = meta["cases"][name]
    E, z, g = make_inputs(c)
    res = vq_oracle.forward_backward_dense(z, E, c["beta"], train_vq=c["train_vq"], g_quantized=g, g_loss=c["g_loss"])
    ref_idx = arrays[f"{name}/idx"].astype(np.int64)
    rows = z.numpy().reshape(-1, c["D"])
    n_mis, gap = near_tie_report(rows, E.numpy(), res.indices.numpy(), ref_idx)
    assert gap <= NEAR_TIE_EXACT and n_mis <= max(1, len(ref_idx) // 500)
    assert abs(float(res.loss) - m["loss"]) <= 1e-6 * abs(m["loss"])
    assert abs(float(res.perplexity) - m["perplexity"]) <= 1e-5 * abs(m["perplexity"])
    assert (res.dE is None) == m["dE_is_none"]
    if n_mis == 0:
        assert abs(float(res.dz.abs().sum()) - m["dz_l1"]) <= 1e-5 * m["dz_l1"]
    # reduced form == dense form (SURVEY.md Appendix A.3 identities)
    comp = vq_oracle.forward_compact(z, E, c["beta"], indices=res.indices)
    assert torch.equal(comp.quantized, res.quantized)
    assert abs(float(comp.loss) - float(res.loss)) <= 1e-6 * abs(float(res.loss))
    assert abs(float(comp.perplexity) - float(res.perplexity)) <= 1e-5 * abs(float(res.perplexity))
    dz, dE = vq_oracle.backward_compact(z, E, res.indices, c["beta"], g, c["g_loss"], c["train_vq"])
    assert rel_err(dz.numpy(), res.dz.numpy()) <= 1e-6
    if res.dE is not None:
        assert rel_err(dE.numpy(), res.dE.numpy()) <= 1e-5


def test_torch_oracle_pinned_to_reference():
    """Bit-for-bit against the real reference class (authoring container only)."""
    if vq_oracle.import_reference_class() is None:
        pytest.skip("/root/reference absent (GPU box): pinned through tests/golden instead")
    assert vq_oracle.check_against_reference(shapes=((2, 4, 3, 8), (4, 64, 50, 256), (2, 128, 100, 512)))


def test_reference_view_errors():
    """vector_quantizer.py:32 `.view(-1, D)`: non-contiguous or indivisible inputs raise RuntimeError."""
    E = vq_oracle.init_codebook(8, 4)
    with pytest.raises(RuntimeError):
        vq_oracle.forward_dense(torch.randn(2, 3, 4).permute(0, 2, 1), E, 0.25)
    with pytest.raises(RuntimeError):
        vq_oracle.forward_dense(torch.randn(2, 5, 3), E, 0.25)


def test_frozen_codebook_semantics():
    """set_train_vq(False): loss value unchanged, no codebook gradient (vector_quantizer.py:47-50)."""
    c = CASES["small_full"]
    E, z, g = make_inputs(c)
    a = vq_oracle.forward_backward_dense(z, E, 0.25, train_vq=True, g_quantized=g)
    b = vq_oracle.forward_backward_dense(z, E, 0.25, train_vq=False, g_quantized=g)
    assert torch.equal(a.loss, b.loss) and torch.equal(a.dz, b.dz) and b.dE is None and a.dE is not None


def test_jitter_oracle_pinned():
    """oracle/jitter_oracle.py vs the reference Jitter (when present) and vs the committed decision fixture."""
    import os
    from oracle import jitter_oracle
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "jitter_golden.npz"))
    np.random.seed(int(g["seed"]))
    src = jitter_oracle.source_columns(int(g["T"]), float(g["p"]))
    assert np.array_equal(src, g["src"].astype(np.int64))
    assert 0.7 < (src != np.arange(len(src))).mean() < 0.8       # jitter.py:50 replaces with probability 1 - p
    if os.path.isdir("/root/reference"):
        assert jitter_oracle.check_against_reference() and jitter_oracle.check_against_reference(5, (2, 8, 500), 0.12)
    # the product's host-side draw consumes np.random identically
    from importlib import import_module
    jit = import_module("acoustic_locating_vq-vae_b200.jitter")
    np.random.seed(int(g["seed"]))
    assert np.array_equal(jit.draw_source_columns(int(g["T"]), float(g["p"])), g["src"].astype(np.int32))


def test_jitter_single_column_raises_like_the_reference():
    """jitter.py:57,68: with T == 1 a drawn replacement indexes column 1 of a one-column tensor -> IndexError; the host-side
    decision draw must fail the same way (and not hand the kernel a source column outside the row)."""
    import numpy as np
    from importlib import import_module
    jit = import_module("acoustic_locating_vq-vae_b200.jitter")
    np.random.seed(0)
    with pytest.raises(IndexError):
        for _ in range(200):                       # the reference replaces with probability 1 - p (jitter.py:50)
            jit.draw_source_columns(1, 0.001)
    np.random.seed(0)
    assert jit.draw_source_columns(1, 1.0).tolist() == [0]      # no replacement drawn: untouched, like the reference
