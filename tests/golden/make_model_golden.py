"""Generates tests/golden/model_golden.npz from the REAL reference model classes (run in the authoring container:
/root/reference present, CPU).  The reference cannot travel to the GPU box (pip builds an empty wheel: pyproject.toml
names the package `acustic_...` while the directory is `acoustic_...`, and the scripts need datasets), so the model-level
configurations of BASELINE.json are pinned at the quantizer's boundary instead:

  configs[1] 'rir_train'   ConvolutionalVQVAE exactly as train_rir.py:124-149 builds it, one full training step
                           (recon + vq loss, backward, Adam 1e-3): what `_vq` received (z), what it returned, the
                           gradients that reached it (g_quantized, g_loss), what it sent back (dz), dE, and the
                           codebook after the Adam step.
  configs[2] 'echoed'      EchoedSpeechReconModel as train_echoed_speech.py:21-27,45-46 with set_train_encoder(True)
                           (encoder_training_echoed_model.py:43-46: the only script where gradient flows through the
                           frozen quantizers): both quantizers' z, outputs, g_quantized and dz (pure straight-through).
  configs[4] 'location'    the RIR-side quantizer of train_location.py:63-75 (frozen, eval): z, indices, perplexity, and
                           the (B, 201, K) one-hot view's column sums.
Batches are cut to 2-4 items so the fixture stays small; widths, K, D, T are the scripts' own.

    python tests/golden/make_model_golden.py
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
sys.path[:0] = [REF, os.path.join(REF, "src")]
from acoustic_locating_vq_vae.vq_vae.convolutional_vq_vae import ConvolutionalVQVAE          # noqa: E402
from acoustic_locating_vq_vae.vq_vae.echoed_speech_model import EchoedSpeechReconModel       # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "model_golden.npz")
rec = {}


class Tap(torch.nn.Module):
    """Wraps the reference quantizer: records what crosses its boundary (values now, gradients via hooks)."""

    def __init__(self, vq, name):
        super().__init__()
        self.vq, self.name = vq, name

    def get_embedding_dim(self):
        return self.vq.get_embedding_dim()

    def set_train_vq(self, flag):
        self.vq.set_train_vq(flag)

    def forward(self, z):
        n = self.name
        rec[f"{n}/E"] = self.vq._embedding.weight.detach().numpy().copy()
        rec[f"{n}/z"] = z.detach().numpy().copy()
        if z.requires_grad:
            z.register_hook(lambda g: rec.__setitem__(f"{n}/dz", g.detach().numpy().copy()))
        loss, q, perp, enc = self.vq(z)
        rec[f"{n}/loss"] = np.float32(loss.item())
        rec[f"{n}/q"] = q.detach().numpy().copy()
        rec[f"{n}/perplexity"] = np.float32(perp.item())
        rec[f"{n}/idx"] = enc.argmax(1).numpy().astype(np.int32)
        rec[f"{n}/train_vq"] = np.int32(bool(self.vq._train_vq))
        if q.requires_grad:
            q.register_hook(lambda g: rec.__setitem__(f"{n}/g_q", g.detach().numpy().copy()))
        if loss.requires_grad:
            loss.register_hook(lambda g: rec.__setitem__(f"{n}/g_loss", np.float32(g.item())))
        return loss, q, perp, enc


def standardise(x):
    return (x - x.mean(dim=(1, 2), keepdim=True)) / x.std(dim=(1, 2), keepdim=True)


def rir_train():
    # train_rir.py:121-149: in_channels 500, hidden 1024, 2 residual layers of 64, D = 64, K = 1024, beta 0.25, no jitter, 1 output channel
    torch.manual_seed(11); np.random.seed(11)
    model = ConvolutionalVQVAE(in_channels=500, num_hiddens=1024, embedding_dim=64, num_residual_layers=2, num_residual_hiddens=64,
                               commitment_cost=0.25, num_embeddings=1024, use_jitter=False, out_channels=1)
    model._vq = Tap(model._vq, "rir_train")
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    model.train()
    x = standardise(torch.randn(4, 201, 500)).permute(0, 2, 1).contiguous()       # train_rir.py:42-45: frames become channels
    target = torch.randn(4, 1, 201)
    vq_loss, recon, perp = model(x)
    loss = torch.nn.functional.mse_loss(recon, target) + vq_loss                  # train_rir.py:54-58
    opt.zero_grad()
    loss.backward()
    rec["rir_train/dE"] = model._vq.vq._embedding.weight.grad.numpy().copy()
    opt.step()
    rec["rir_train/E_after_adam"] = model._vq.vq._embedding.weight.detach().numpy().copy()


def echoed():
    torch.manual_seed(12); np.random.seed(12)
    speech = ConvolutionalVQVAE(201, 1024, 128, 3, 1024, 0.25, 1024)              # train_speech.py:24-41
    rir = ConvolutionalVQVAE(500, 1024, 64, 2, 64, 0.25, 1024, use_jitter=False, out_channels=1)
    for m in (speech, rir):
        m._vq._embedding.weight.data.normal_()                                    # "trained-like" codebooks
    model = EchoedSpeechReconModel(rir, speech, 201, 1024, 2, 1024, True)         # train_echoed_speech.py:21-27,45-46
    model.set_train_encoder(True)                                                 # encoder_training_echoed_model.py:44
    speech._vq = Tap(speech._vq, "echoed_speech")
    rir._vq = Tap(rir._vq, "echoed_rir")
    model.train()
    x = standardise(torch.randn(2, 201, 500))
    recon, sp, rp = model(x, x.permute(0, 2, 1).contiguous())
    torch.nn.functional.mse_loss(recon, x).backward()                             # train_echoed_speech.py:89-90


def location():
    torch.manual_seed(13); np.random.seed(13)
    rir = ConvolutionalVQVAE(500, 1024, 64, 2, 64, 0.25, 1024, use_jitter=False, out_channels=1)
    rir._vq._embedding.weight.data.normal_()
    rir._vq.set_train_vq(False)
    tap = Tap(rir._vq, "location_rir")
    rir._vq = tap
    rir.eval()
    x = standardise(torch.randn(4, 201, 500)).permute(0, 2, 1).contiguous()       # train_location.py:63-66
    _, _, _, enc = rir.get_latent_representation(x)                               # :69
    enc3 = enc.reshape(4, 201, 1024)                                              # :74
    rec["location_rir/enc_colsum"] = enc3.sum(1).numpy().astype(np.float32)       # (B, K): usage per sample


if __name__ == "__main__":
    rir_train()
    echoed()
    location()
    np.savez_compressed(OUT, **rec)
    print(OUT, f"{os.path.getsize(OUT) / 1e6:.1f} MB")
    for k in sorted(rec):
        v = rec[k]
        print(f"  {k:32s} {getattr(v, 'shape', ())} {v.dtype}")
