"""The drop-in quantizer at the boundary of the REAL reference models (BASELINE.json configs[1], [2], [4]).

tests/golden/model_golden.npz was recorded by tests/golden/make_model_golden.py from the unmodified reference classes
(ConvolutionalVQVAE as train_rir.py builds it, EchoedSpeechReconModel with set_train_encoder(True), the RIR-side
quantizer of train_location.py): for every quantizer call what it received (z, codebook), what it returned, the
gradients that reached it from the decoder / the loss and what it sent back -- plus the codebook after the script's
Adam step.  Here the B200 module gets the same z and the same upstream gradients and must reproduce all of it.
(The reference cannot run on the GPU box: its package does not install)
"""
import os

import numpy as np
import pytest
import torch

from util import NEAR_TIE_EXACT, near_tie_report, rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = 1e-5


@pytest.fixture(scope="module")
def G():
    return np.load(os.path.join(ROOT, "tests", "golden", "model_golden.npz"))


def _run(G, name, exact):
    import b200vq
    dev = torch.device("cuda:0")
    E = torch.from_numpy(G[f"{name}/E"])
    K, D = E.shape
    vq = b200vq.VectorQuantizer(K, D, 0.25, exact=exact).to(dev)
    vq._embedding.weight.data.copy_(E)
    vq.set_train_vq(bool(G[f"{name}/train_vq"]))
    z = torch.from_numpy(G[f"{name}/z"]).to(dev).requires_grad_(True)
    loss, q, perp, enc = vq(z)
    return vq, z, loss, q, perp, enc


@pytest.mark.parametrize("exact", [False, True])
@pytest.mark.parametrize("name", ["rir_train", "echoed_speech", "echoed_rir", "location_rir"])
def test_forward_matches_the_reference_models_quantizer(G, name, exact):
    vq, z, loss, q, perp, enc = _run(G, name, exact)
    idx = vq.last_indices.cpu().numpy()
    ref_idx = G[f"{name}/idx"]
    D = z.shape[1]
    n_mis, gap = near_tie_report(G[f"{name}/z"].reshape(-1, D), G[f"{name}/E"], idx, ref_idx)
    assert n_mis <= 3 and gap <= NEAR_TIE_EXACT, f"{name}: {n_mis} rows differ from the reference, worst fp64 gap {gap:.2e}"
    assert torch.equal(enc.argmax(1).int().cpu(), torch.from_numpy(idx)) and float(enc.sum()) == idx.size
    same = idx == ref_idx
    qr = q.detach().cpu().numpy().reshape(-1, D)
    assert rel_err(qr[same], G[f"{name}/q"].reshape(-1, D)[same]) <= RTOL
    tol = RTOL + 2.0 * n_mis / idx.size            # a differing near-tie row moves the means by at most ~1/N
    assert abs(float(loss) - float(G[f"{name}/loss"])) <= tol * abs(float(G[f"{name}/loss"]))
    assert abs(float(perp) - float(G[f"{name}/perplexity"])) <= tol * float(G[f"{name}/perplexity"]) + 1e-3 * n_mis


@pytest.mark.parametrize("name", ["rir_train", "echoed_speech", "echoed_rir"])
def test_backward_and_adam_step_match_the_reference_models(G, name):
    vq, z, loss, q, perp, enc = _run(G, name, exact=False)
    D = z.shape[1]
    dev = z.device
    g_q = torch.from_numpy(G[f"{name}/g_q"]).to(dev)
    opt = torch.optim.Adam(vq.parameters(), lr=1e-3)              # train_rir.py:151 / train_echoed_speech.py
    opt.zero_grad()
    if f"{name}/g_loss" in G.files:
        torch.autograd.backward([loss, q], [torch.tensor(float(G[f"{name}/g_loss"]), device=dev), g_q])
    else:
        torch.autograd.backward([q], [g_q])                       # the echoed model drops the quantizers' loss
    idx = vq.last_indices.cpu().numpy()
    same = idx == G[f"{name}/idx"]
    dz = z.grad.cpu().numpy().reshape(-1, D)
    assert rel_err(dz[same], G[f"{name}/dz"].reshape(-1, D)[same]) <= RTOL
    if not bool(G[f"{name}/train_vq"]):
        assert vq._embedding.weight.grad is None                  # frozen codebook: pure straight-through
        assert np.array_equal(dz[same], G[f"{name}/g_q"].reshape(-1, D)[same])
        return
    if same.all():
        assert rel_err(vq._embedding.weight.grad.cpu().numpy(), G[f"{name}/dE"]) <= RTOL
        opt.step()
        # Adam's first step moves every touched weight by lr * sign(g) (up to eps): compare the codebooks themselves
        assert rel_err(vq._embedding.weight.detach().cpu().numpy(), G[f"{name}/E_after_adam"]) <= RTOL


def test_location_onehot_view_matches_the_reference(G):
    """train_location.py:74: encodings.reshape(B, 201, K) feeds the location head; its per-sample code usage must
    equal the reference's, and the index path (OneHotLinear) must see the same codes."""
    vq, z, loss, q, perp, enc = _run(G, "location_rir", exact=False)
    B = z.shape[0]
    idx = vq.last_indices.cpu().numpy()
    if np.array_equal(idx, G["location_rir/idx"]):
        assert np.array_equal(enc.reshape(B, 201, -1).sum(1).cpu().numpy(), G["location_rir/enc_colsum"])
    assert torch.equal(enc.reshape(B, 201, -1).argmax(2).int().cpu(), torch.from_numpy(idx).view(B, 201))
