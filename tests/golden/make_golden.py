"""Generate golden fixtures from the REAL reference VectorQuantizer (authoring container only).

Run:  python tests/golden/make_golden.py
Needs /root/reference (read-only); imports
`acoustic_locating_vq_vae.vq_vae.vector_quantizer.VectorQuantizer` unmodified, runs it on CPU
(torch fp32, MKL) on seeded inputs and writes `tests/golden/vq_golden.npz` + `vq_golden.json`.
The fixtures travel to the GPU box; /root/reference does not.

Input recipe (same as SURVEY.md Appendix A.2, reproduced by tests/cases.py without the reference):
    torch.manual_seed(seed); vq = VectorQuantizer(K, D, 0.25)      # nn.Embedding N(0,1) draw, then U(-1/K,1/K)
    [vq._embedding.weight.data.normal_()]                          # init == "normal"
    z = torch.randn(B, D, T); g = torch.randn(B, D, T)
    loss, q, perp, enc = vq(z); (g_loss*loss + (g*q).sum()).backward()
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from cases import CASES, make_inputs  # noqa: E402
from oracle.vq_oracle import import_reference_class  # noqa: E402


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.detach().contiguous().numpy().tobytes()).hexdigest()[:16]


def main():
    Ref = import_reference_class()
    assert Ref is not None, "needs /root/reference"
    torch.set_num_threads(8)
    arrays = {}
    meta = {"torch": torch.__version__, "generator": "tests/golden/make_golden.py",
            "reference": "guy3540/Acoustic_Locating_VQ-VAE src/acoustic_locating_vq_vae/vq_vae/vector_quantizer.py",
            "cases": {}}
    for name, c in CASES.items():
        # Build through the reference ctor so the RNG consumption is the reference's own, and
        # check that cases.make_inputs reproduces it bit for bit without the reference class.
        torch.manual_seed(c["seed"])
        vq = Ref(c["K"], c["D"], c["beta"])
        if c["init"] in ("normal", "data"):
            vq._embedding.weight.data.normal_()
        vq.set_train_vq(c["train_vq"])
        z = torch.randn(*c["shape"])
        g = torch.randn(*c["shape"])
        E2, z2, g2 = make_inputs(c)
        assert torch.equal(z2, z) and torch.equal(g2, g), name
        if c["init"] == "data":
            vq._embedding.weight.data.copy_(E2)
        assert torch.equal(E2, vq._embedding.weight.detach()), name
        z.requires_grad_(True)
        loss, q, perp, enc = vq(z)
        (c["g_loss"] * loss + (g * q).sum()).backward()
        idx = enc.argmax(dim=1).to(torch.int32)
        assert torch.equal(enc.sum(1), torch.ones(enc.shape[0]))
        dE = vq._embedding.weight.grad
        N = enc.shape[0]
        m = {"N": N, "loss": float(loss), "perplexity": float(perp), "idx_sum": int(idx.sum()),
             "idx_sha": hashlib.sha256(idx.numpy().tobytes()).hexdigest()[:16],
             "n_unique": int(idx.unique().numel()), "z_sha": sha(z), "E_sha": sha(vq._embedding.weight),
             "q_sha": sha(q), "dz_l1": float(z.grad.abs().sum()), "q_l1": float(q.detach().abs().sum()),
             "dE_l1": None if dE is None else float(dE.abs().sum()), "dE_is_none": dE is None}
        meta["cases"][name] = m
        arrays[f"{name}/idx"] = idx.numpy().astype(np.int16 if c["K"] <= 32767 else np.int32)
        if c["store_full"]:
            arrays[f"{name}/z"] = z.detach().numpy()
            arrays[f"{name}/E"] = vq._embedding.weight.detach().numpy()
            arrays[f"{name}/g"] = g.numpy()
            arrays[f"{name}/q"] = q.detach().numpy()
            arrays[f"{name}/dz"] = z.grad.numpy()
            if dE is not None:
                arrays[f"{name}/dE"] = dE.numpy()
        else:
            # row samples + codebook-gradient row sums keep the fixture small
            D = c["D"]
            rows = np.linspace(0, N - 1, 64).astype(np.int64)
            arrays[f"{name}/rows"] = rows
            arrays[f"{name}/q_rows"] = q.detach().view(-1, D)[rows].numpy()
            arrays[f"{name}/dz_rows"] = z.grad.view(-1, D)[rows].numpy()
            if dE is not None:
                arrays[f"{name}/dE_rowsum"] = dE.sum(1).numpy()
                arrays[f"{name}/dE_colsum"] = dE.sum(0).numpy()
        print(f"{name:28s} N={N:6d} loss={m['loss']:.8f} perp={m['perplexity']:.6f} idx_sha={m['idx_sha']} uniq={m['n_unique']}")
    np.savez_compressed(os.path.join(HERE, "vq_golden.npz"), **arrays)
    with open(os.path.join(HERE, "vq_golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote", os.path.getsize(os.path.join(HERE, "vq_golden.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
