"""GPU integration tests: the quantizer inside the callers' graphs (BASELINE.json configs[2] and configs[4]).

The reference's model classes cannot travel to the GPU box and its conv stacks are out of scope, so the callers
are re-expressed here as small stand-ins with the same dataflow and shapes:
  * `_ConvVQVAE`      convolutional_vq_vae.py:93-100   encoder conv -> _pre_vq_conv -> _vq -> decoder
  * jitter            modules/jitter.py:47-70          in-place edit of the quantizer's output before the decoder
  * echoed model      echoed_speech_model.py:36-56     two frozen quantizers, pad + concat, (un)detached
  * location model    train_location.py:69-75          encodings.reshape(B, T, K) -> MLP
Both the B200 quantizer and the torch oracle quantizer (oracle/vq_oracle.py, same aten ops as the reference) are
fed the SAME z (cudnn TF32 convs must not be re-run under different flags), and everything
downstream -- losses, reconstructions, gradients, Adam-updated weights -- must agree to 1e-5 relative.
"""
import copy

import numpy as np

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import vq_oracle

pytestmark = pytest.mark.gpu
RTOL = 1e-5


class _OracleVQ(nn.Module):
    """Reference-equivalent quantizer (torch ops) with the reference's surface."""

    def __init__(self, K, D, beta):
        super().__init__()
        self._embedding = nn.Embedding(K, D)
        self._embedding.weight.data.uniform_(-1 / K, 1 / K)
        self._commitment_cost = beta
        self._train_vq = True

    def set_train_vq(self, flag):
        self._train_vq = flag

    def forward(self, inputs):
        r = vq_oracle.forward_dense(inputs, self._embedding.weight, self._commitment_cost, self._train_vq)
        return r.loss, r.quantized, r.perplexity, r.encodings


class _ConvVQVAE(nn.Module):
    def __init__(self, in_ch, hidden, D, vq):
        super().__init__()
        self._encoder = nn.Conv1d(in_ch, hidden, 3, padding=1)
        self._pre_vq_conv = nn.Conv1d(hidden, D, 3, padding=1)      # convolutional_vq_vae.py:32-38
        self._vq = vq
        self._decoder = nn.Conv1d(D, in_ch, 3, padding=1)

    def latent(self, x):
        return self._pre_vq_conv(F.relu(self._encoder(x)))

    def forward(self, x, jitter_cols=None):
        z = self.latent(x)
        loss, q, perp, _ = self._vq(z)                                # :98
        if jitter_cols is not None:                                   # jitter.py:53-68: IN-PLACE on the VQ output
            orig = q.detach().clone()
            for i, j in jitter_cols:
                q[:, :, i] = orig[:, :, j]
        return loss, self._decoder(q), perp


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / max(float(b.double().abs().max()), 1e-30))


def _pair(K, D, beta, dev, seed, **kw):
    import b200vq
    torch.manual_seed(seed)
    ref = _OracleVQ(K, D, beta).to(dev)
    ref._embedding.weight.data.normal_()
    mine = b200vq.VectorQuantizer(K, D, beta, **kw).to(dev)
    mine._embedding.weight.data.copy_(ref._embedding.weight.data)
    return mine, ref


@pytest.mark.parametrize("shape,K,jitter", [((32, 201, 500, 128), 1024, True),     # train_speech.py: B,in,T,D
                                             ((16, 500, 201, 64), 1024, False)])   # train_rir.py (no jitter)
def test_training_step_inside_conv_stack(shape, K, jitter):
    """One optimiser step of an encoder -> quantizer -> decoder model (configs[0]/[1] shapes): same z into both
    quantizers, then loss, reconstruction, every gradient and the Adam-updated weights must agree."""
    dev = torch.device("cuda:0")
    B, C, T, D = shape
    mine, ref = _pair(K, D, 0.25, dev, 0)
    torch.manual_seed(1)
    model = _ConvVQVAE(C, 64, D, mine).to(dev)
    twin = copy.deepcopy(model)
    twin._vq = ref
    x = torch.randn(B, C, T, device=dev)
    cols = [(0, 1), (7, 6), (T - 1, T - 2), (100, 101)] if jitter else None
    outs = []
    for m in (model, twin):
        opt = torch.optim.Adam(m.parameters(), lr=1e-3)
        z = model.latent(x).detach().requires_grad_(True)             # the SAME z for both quantizers
        loss, q, perp, enc = m._vq(z)
        if cols is not None:
            orig = q.detach().clone()
            for i, j in cols:
                q[:, :, i] = orig[:, :, j]                            # in-place, like Jitter
        recon = m._decoder(q)
        total = F.mse_loss(recon, x) + loss                           # train_speech.py:74,88
        opt.zero_grad()
        total.backward()
        opt.step()
        outs.append(dict(loss=loss.detach(), perp=perp, total=total.detach(), dz=z.grad, enc=enc,
                         dE=m._vq._embedding.weight.grad, E=m._vq._embedding.weight.detach(),
                         dec=m._decoder.weight.detach(), recon=recon.detach()))
    a, b = outs
    assert torch.equal(a["enc"], b["enc"])                            # identical codes
    for k in ("loss", "perp", "total", "dz", "dE", "E", "dec", "recon"):
        assert _rel(a[k], b[k]) <= RTOL, (k, _rel(a[k], b[k]))


def test_echoed_model_dataflow():
    """echoed_speech_model.py:36-56: both quantizers frozen (`set_train_vq(False)`), RIR latent padded on time and
    concatenated on channels, detached unless flag_train_encoder; encoder_training_echoed_model.py:44 flips it."""
    dev = torch.device("cuda:0")
    B = 8
    sp_m, sp_r = _pair(1024, 128, 0.25, dev, 2)
    rir_m, rir_r = _pair(1024, 64, 0.25, dev, 3)
    for v in (sp_m, sp_r, rir_m, rir_r):
        v.set_train_vq(False)
    torch.manual_seed(4)
    dec = nn.Conv1d(192, 201, 3, padding=1).to(dev)
    z_sp = torch.randn(B, 128, 500, device=dev)
    z_rir = torch.randn(B, 64, 201, device=dev)
    x = torch.randn(B, 201, 500, device=dev)
    for train_encoder in (False, True):
        res = []
        for sp, rir in ((sp_m, rir_m), (sp_r, rir_r)):
            a = z_sp.clone().requires_grad_(True)
            b = z_rir.clone().requires_grad_(True)
            _, q_rir, p_rir, _ = rir(b)
            _, q_sp, p_sp, _ = sp(a)
            q_rir = F.pad(q_rir, (0, q_sp.size(2) - q_rir.size(2)))   # :41-49
            cat = torch.cat((q_sp, q_rir), 1) if train_encoder else torch.cat((q_sp.detach(), q_rir.detach()), 1)
            loss = F.mse_loss(dec(cat), x)
            dec.zero_grad()
            loss.backward()
            res.append(dict(loss=loss.detach(), p=torch.stack((p_sp, p_rir)), g=dec.weight.grad.clone(),
                            da=a.grad, db=b.grad, dE=sp._embedding.weight.grad))
        m, r = res
        assert _rel(m["loss"], r["loss"]) <= RTOL and _rel(m["p"], r["p"]) <= RTOL and _rel(m["g"], r["g"]) <= RTOL
        assert m["dE"] is None and r["dE"] is None                    # frozen codebooks never get a gradient
        if train_encoder:                                             # pure straight-through: dz == upstream gradient
            assert _rel(m["da"], r["da"]) <= RTOL and _rel(m["db"], r["db"]) <= RTOL
        else:
            assert m["da"] is None and r["da"] is None


def test_location_model_consumes_encodings():
    """train_location.py:69-75: the RIR quantizer's one-hot `encodings` reshaped to (B, T, K) feeds the MLP."""
    dev = torch.device("cuda:0")
    B, T, K, D = 16, 201, 1024, 64
    mine, ref = _pair(K, D, 0.25, dev, 5)
    mine.set_train_vq(False)
    ref.set_train_vq(False)
    torch.manual_seed(6)
    fc1 = nn.Linear(T * K, 64).to(dev)                                # location_model.py:10 (1024 wide there)
    z = torch.randn(B, D, T, device=dev)
    target = torch.rand(B, 1, device=dev)
    outs = []
    for vq in (mine, ref):
        _, _, _, enc = vq(z)
        feat = enc.reshape(B, T, K)                                   # :74
        y = fc1(torch.flatten(feat, start_dim=1)).mean(1, keepdim=True)
        loss = F.mse_loss(y, target)
        fc1.zero_grad()
        loss.backward()
        outs.append((loss.detach(), fc1.weight.grad.clone(), enc))
    assert torch.equal(outs[0][2], outs[1][2])
    assert _rel(outs[0][0], outs[1][0]) <= RTOL and _rel(outs[0][1], outs[1][1]) <= RTOL
    # the same features without the dense one-hot: indices + gather
    idx = mine.last_indices.view(B, T).long()
    cols = (torch.arange(T, device=dev) * K)[None, :] + idx          # column t*K + idx[b,t] of fc1.weight
    y2 = fc1.weight[:, cols].sum(-1).t() + fc1.bias                   # == fc1(one_hot) exactly up to summation order
    y1 = fc1(torch.flatten(outs[0][2].reshape(B, T, K), start_dim=1))
    assert _rel(y2, y1) <= 1e-5


@pytest.mark.parametrize("B,T,K,O,bias", [(16, 201, 1024, 1024, True),      # train_location.py / location_model.py shapes
                                           (5, 20, 64, 32, True), (3, 7, 37, 8, False)])
def test_onehot_linear_equals_linear_on_onehot(B, T, K, O, bias):
    """`fc_1(flatten(one_hot))` (location_model.py:21) == gather of T weight rows per sample."""
    import b200vq
    dev = torch.device("cuda:0")
    torch.manual_seed(7)
    lin = nn.Linear(T * K, O, bias=bias).to(dev)
    idx = torch.randint(0, K, (B, T), device=dev)
    onehot = F.one_hot(idx, K).float()                               # (B, T, K) as train_location.py:74 builds it
    y_ref = lin(torch.flatten(onehot, start_dim=1))
    g = torch.randn(B, O, device=dev)
    gw_ref, = torch.autograd.grad(y_ref, lin.weight, g, retain_graph=bias)
    for sparse in (True, False):
        m = b200vq.OneHotLinear.from_linear(lin, T, K, sparse_grad=sparse)
        y = m(idx.int())
        assert _rel(y, y_ref) <= 1e-5
        y.backward(g)
        gw = m.weight_t.grad
        assert gw.is_sparse == sparse
        gw = gw.to_dense() if sparse else gw
        assert _rel(gw.t(), gw_ref) <= 1e-6
        if bias:
            assert _rel(m.bias.grad, g.sum(0)) <= 1e-6
    assert torch.equal(m.to_linear().weight, lin.weight)
    # straight from the quantizer: no dense one-hot is ever built
    if K % 256 == 0:
        vq = b200vq.VectorQuantizer(K, 64, 0.25, return_encodings=False).to(dev)
        z = torch.randn(B, 64, T, device=dev)
        _, _, _, enc = vq(z)
        assert enc is None
        feats = m(vq.last_indices.view(B, T))
        ref = lin(torch.flatten(F.one_hot(vq.last_indices.view(B, T).long(), K).float(), start_dim=1))
        assert _rel(feats, ref) <= 1e-5


@pytest.mark.parametrize("shape,p", [((32, 128, 500), 0.25), ((4, 64, 201), 0.12), ((2, 3, 2), 0.25), ((1, 1, 1500), 0.5)])
def test_jitter_matches_reference_semantics(shape, p):
    """same np.random seed -> same tensor as modules/jitter.py:47-70, mutated in place; replaced
    columns carry no gradient."""
    import numpy as np
    import b200vq
    from oracle import jitter_oracle
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    x = torch.randn(*shape)
    np.random.seed(99)
    ref = jitter_oracle.jitter(x.clone(), p)
    src = None
    np.random.seed(99)
    src = jitter_oracle.source_columns(shape[2], p)
    z = x.to(dev).requires_grad_(True)
    q = z * 1.0                                      # a non-leaf, like the quantizer's output
    np.random.seed(99)
    out = b200vq.Jitter(p)(q)
    assert out.data_ptr() == q.data_ptr()            # in place, like the reference
    assert torch.equal(out.detach().cpu(), ref)
    g = torch.randn(*shape, device=dev)
    out.backward(g)
    keep = torch.from_numpy(src == np.arange(shape[2])).to(dev)
    assert torch.equal(z.grad, g * keep)


def test_module_is_cuda_graph_capturable():
    """The whole forward + backward of the module (prepare, fused forward, backward; no host sync, no allocation
    outside the caching allocator) can be captured in a CUDA graph and replayed on new data."""
    import b200vq
    dev = torch.device("cuda:0")
    B, D, T, K = 32, 64, 201, 1024
    torch.manual_seed(8)
    vq = b200vq.VectorQuantizer(K, D, 0.25).to(dev)
    vq._embedding.weight.data.normal_()
    static_z = torch.randn(B, D, T, device=dev, requires_grad=True)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            static_z.grad = None
            vq._embedding.weight.grad = None
            loss, q, perp, enc = vq(static_z)
            (loss + q.sum()).backward()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    static_z.grad = None
    vq._embedding.weight.grad = None
    with torch.cuda.graph(g):
        loss, q, perp, enc = vq(static_z)
        (loss + q.sum()).backward()
    new = torch.randn(B, D, T, device=dev)
    static_z.data.copy_(new)
    g.replay()
    torch.cuda.synchronize()
    eager = b200vq.VectorQuantizer(K, D, 0.25).to(dev)
    eager._embedding.weight.data.copy_(vq._embedding.weight.data)
    z2 = new.clone().requires_grad_(True)
    l2, q2, p2, e2 = eager(z2)
    (l2 + q2.sum()).backward()
    assert torch.equal(enc, e2) and torch.equal(q.detach(), q2.detach())
    assert _rel(loss.detach(), l2.detach()) <= 1e-6 and _rel(perp, p2) <= 1e-6
    assert _rel(static_z.grad, z2.grad) <= 1e-6 and _rel(vq._embedding.weight.grad, eager._embedding.weight.grad) <= RTOL


@pytest.mark.parametrize("B,D,T,K", [(32, 128, 500, 1024), (8, 64, 201, 256), (5, 32, 7, 64), (3, 10, 13, 37)])
def test_time_mean_variant_in_front_of_the_quantizer(B, D, T, K):
    """convolutional_vq_vae.py:96-98 with encoder_average_pooling=True: z = mean over time, then the quantizer on B rows.
    The fused call must equal the two reference ops: the pooled z within fp32 summation-order rounding of torch.mean,
    and -- GIVEN that z -- indices bit-exact vs the oracle, outputs and gradients within the north-star tolerance."""
    import b200vq
    from oracle import c_oracle
    dev = torch.device("cuda:0")
    torch.manual_seed(B + T)
    vq = b200vq.VectorQuantizer(K, D, 0.25).to(dev)
    vq._embedding.weight.data.normal_()
    x = torch.randn(B, D, T, device=dev, requires_grad=True)
    z = b200vq.time_mean(x)
    assert z.shape == (B, D, 1)
    ref_z = torch.mean(x.detach(), dim=2, keepdim=True)
    assert float((z.detach() - ref_z).abs().max()) <= 4e-7 * max(1.0, float(ref_z.abs().max()))
    loss, q, perp, enc = b200vq.time_mean_quantize(vq, x)
    rows = z.detach().reshape(B, D).cpu().numpy()
    E = vq._embedding.weight.detach().cpu().numpy()
    idx = c_oracle.argmin(rows, E)
    assert np.array_equal(vq.last_indices.cpu().numpy(), idx)
    fwd = c_oracle.quantize(rows, E, idx, 0.25)
    assert abs(float(loss) - fwd["loss"]) <= RTOL * fwd["loss"] and abs(float(perp) - fwd["perplexity"]) <= RTOL * fwd["perplexity"]
    g = torch.randn(B, D, 1, device=dev)
    (loss + (g * q).sum()).backward()
    dz_ref, dE_ref = c_oracle.backward(g.reshape(B, D).cpu().numpy(), 1.0, rows, E, idx, 0.25, True)
    dx_ref = np.repeat(dz_ref.reshape(B, D, 1), T, axis=2) / T                       # d mean / d x = 1 / T
    assert _rel(x.grad, torch.from_numpy(dx_ref).to(dev)) <= RTOL
    assert _rel(vq._embedding.weight.grad, torch.from_numpy(dE_ref).to(dev)) <= RTOL


def test_module_on_a_second_device_in_the_same_process():
    """The kernels' shared-memory opt-in (cudaFuncSetAttribute) is per device: a module on cuda:1 after one on cuda:0 in the
    same process must launch and give the same indices on the same data.  Needs two visible GPUs."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible GPUs")
    import b200vq
    torch.manual_seed(3)
    E = torch.randn(1024, 64)
    z = torch.randn(8, 64, 201)
    outs = []
    for dev in ("cuda:0", "cuda:1", "cuda:0"):
        vq = b200vq.VectorQuantizer(1024, 64, 0.25).to(dev)
        vq._embedding.weight.data.copy_(E)
        zz = z.to(dev).requires_grad_(True)
        loss, q, perp, enc = vq(zz)
        (loss + q.sum()).backward()
        torch.cuda.synchronize(dev)
        outs.append((vq.last_indices.cpu(), float(loss), zz.grad.cpu(), vq._embedding.weight.grad.cpu()))
    for o in outs[1:]:
        assert torch.equal(o[0], outs[0][0]) and abs(o[1] - outs[0][1]) <= 1e-6 * abs(outs[0][1])
        assert torch.equal(o[2], outs[0][2]) and torch.allclose(o[3], outs[0][3], rtol=1e-5, atol=1e-9)
