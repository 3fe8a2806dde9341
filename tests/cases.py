"""Seeded parity cases shared by the golden generator, the CPU tests and the GPU tests.

Shapes follow SURVEY.md section 8(d) / Appendix C: (B, D, T) is what `_pre_vq_conv` hands to the
quantizer (`convolutional_vq_vae.py:95-98`); the quantizer flattens it to N = B*T rows of D
consecutive floats (`vector_quantizer.py:32`).
"""
from __future__ import annotations

import torch


def _case(seed, shape, K, init, train_vq=True, g_loss=1.0, store_full=False, beta=0.25):
    return dict(seed=seed, shape=tuple(shape), K=K, D=shape[1], init=init, train_vq=train_vq,
                g_loss=g_loss, store_full=store_full, beta=beta)


CASES = {
    # BASELINE.json configs[0]: speech VQ, train_speech.py defaults (B=32, D=128, T=500, K=1024)
    "speech_uniform":  _case(0, (32, 128, 500), 1024, "uniform"),
    "speech_normal":   _case(0, (32, 128, 500), 1024, "normal"),
    # train_rir.py defaults (B=32, D=64, T=201)
    "rir32_uniform":   _case(0, (32, 64, 201), 1024, "uniform"),
    "rir32_normal":    _case(0, (32, 64, 201), 1024, "normal"),
    # BASELINE.json configs[1]: RIR quantizer, batch 256  (N = 51 456; not a multiple of 128)
    "rir256_uniform":  _case(0, (256, 64, 201), 1024, "uniform"),
    "rir256_normal":   _case(0, (256, 64, 201), 1024, "normal", g_loss=0.7),
    # train_location.py RIR side (B=16), frozen quantizer as echoed_speech_model.py:17-18 sets it
    "loc16_normal":    _case(0, (16, 64, 201), 1024, "normal", train_vq=False),
    "loc16_uniform":   _case(0, (16, 64, 201), 1024, "uniform", train_vq=False),
    # sweep corners at oracle-friendly N (configs[3])
    "sweep_k512_d64":  _case(1, (64, 64, 64), 512, "normal"),
    "sweep_k2048_d256": _case(2, (16, 256, 64), 2048, "normal"),
    "sweep_k8192_d128": _case(3, (8, 128, 96), 8192, "normal"),
    # codebook that fits the data (small q - z: exercises cancellation in loss / dE)
    "fitted_d64":      _case(4, (8, 64, 100), 256, "data", g_loss=1.3),
    # small cases with every tensor stored
    "small_full":      _case(5, (4, 64, 50), 256, "normal", store_full=True),
    "small_frozen":    _case(6, (4, 64, 50), 256, "normal", train_vq=False, store_full=True, g_loss=0.5),
    "small_uniform":   _case(7, (4, 32, 40), 128, "uniform", store_full=True),
    # ragged / edge shapes: D and K that are multiples of nothing, pooled (B, D, 1) variant
    # (convolutional_vq_vae.py:96-97), a single row, K = 1
    "odd_shape":       _case(8, (3, 5, 7), 37, "normal", store_full=True),
    "pooled":          _case(9, (8, 64, 1), 128, "normal", store_full=True),
    "single_row":      _case(10, (1, 16, 1), 16, "normal", store_full=True),
    "one_code":        _case(11, (2, 8, 5), 1, "normal", store_full=True),
    "d96_k384":        _case(12, (4, 96, 33), 384, "normal", store_full=True),
}


def make_inputs(c):
    """Reproduce (codebook, z, g) of a case without the reference class.

    Consumes the global torch RNG exactly like the reference constructor does
    (`vector_quantizer.py:15-16`: nn.Embedding's N(0,1) draw, then the U(-1/K, 1/K) overwrite).
    """
    K, D = c["K"], c["D"]
    torch.manual_seed(c["seed"])
    E = torch.empty(K, D)
    E.normal_()                      # nn.Embedding.reset_parameters
    E.uniform_(-1.0 / K, 1.0 / K)    # vector_quantizer.py:16
    if c["init"] in ("normal", "data"):
        E.normal_()
    z = torch.randn(*c["shape"])
    g = torch.randn(*c["shape"])
    if c["init"] == "data":
        rows = z.view(-1, D)
        sel = torch.arange(K) * (rows.shape[0] // K)
        E = rows[sel] + 0.01 * torch.randn(K, D)
    return E.contiguous(), z, g
