"""CPU tests of the boundary: libb200vq.so loads, exports every symbol include/b200vq.h declares,
refuses to run without a B200, and the host-side module mirrors the reference's surface."""
import ctypes
import os
import pickle
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "b200vq.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vq_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    import b200vq
    names = header_functions()
    assert len(names) >= 15
    raw = ctypes.CDLL(b200vq.SO_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/b200vq.h but not exported"
    from importlib import import_module
    sigs = import_module("acoustic_locating_vq-vae_b200._lib").SIGNATURES
    assert sorted(sigs) == names, "ctypes binding and header disagree"
    assert lib.vq_abi_version() == import_module("acoustic_locating_vq-vae_b200._lib").ABI_VERSION == 18


def test_no_cpu_fallback(lib):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert lib.vq_device_check() == 3            # VQ_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.vq_last_error()
    rc = lib.vq_prepare_codebook(None, 8, 4, None, None, None, None)
    assert rc != 0
    import b200vq
    vq = b200vq.VectorQuantizer(8, 4, 0.25)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vq(torch.randn(2, 4, 3))


def test_path_selection(lib):
    assert lib.vq_forward_uses_tensor_path(51456, 1024, 64, 0) == 1
    assert lib.vq_forward_uses_tensor_path(16000, 1024, 128, 0) == 1
    assert lib.vq_forward_uses_tensor_path(16000, 1024, 128, 4) == 0     # VQ_FLAG_EXACT
    assert lib.vq_forward_uses_tensor_path(21, 37, 5, 0) == 0
    assert lib.vq_workspace_bytes(51456, 1024, 64, 0) >= 51456 * 8


def test_module_surface_matches_reference():
    """Constructor, attributes, accessors and state-dict key of vector_quantizer.py:9-27."""
    import b200vq
    torch.manual_seed(3)
    vq = b200vq.VectorQuantizer(num_embeddings=32, embedding_dim=8, commitment_cost=0.25)
    assert list(vq.state_dict().keys()) == ["_embedding.weight"]
    assert vq._embedding.weight.shape == (32, 8) and vq._embedding.weight.abs().max() <= 1 / 32
    assert vq.get_embedding_dim() == 8 and vq._num_embeddings == 32 and vq._commitment_cost == 0.25
    assert vq._train_vq is True and vq._flag_flatten is True
    vq.set_train_vq(False)
    assert vq._train_vq is False
    # same RNG consumption as the reference constructor (nn.Embedding draw, then uniform_)
    torch.manual_seed(3)
    w = torch.empty(32, 8).normal_()
    w.uniform_(-1 / 32, 1 / 32)
    assert torch.equal(w, vq._embedding.weight.detach())
    vq2 = pickle.loads(pickle.dumps(vq))
    assert torch.equal(vq2._embedding.weight, vq._embedding.weight) and vq2._train_vq is False


def test_swap_quantizers_on_reference_models():
    from oracle.vq_oracle import import_reference_class
    import b200vq
    Ref = import_reference_class()
    if Ref is None:
        pytest.skip("/root/reference absent")
    from acoustic_locating_vq_vae.vq_vae.convolutional_vq_vae import ConvolutionalVQVAE
    from acoustic_locating_vq_vae.vq_vae.echoed_speech_model import EchoedSpeechReconModel
    rir = ConvolutionalVQVAE(20, 16, 8, 1, 8, 0.25, 32, use_jitter=False, out_channels=1)
    speech = ConvolutionalVQVAE(12, 16, 8, 1, 8, 0.25, 32)
    model = EchoedSpeechReconModel(rir, speech, 12, 16, 1, 8, True)
    w_before = model.rir_model._vq._embedding.weight.detach().clone()
    assert b200vq.swap_quantizers(model) == 2
    assert isinstance(model.rir_model._vq, b200vq.VectorQuantizer)
    assert isinstance(model.speech_model._vq, b200vq.VectorQuantizer)
    assert model.rir_model._vq._train_vq is False            # echoed_speech_model.py:17-18 froze them
    assert torch.equal(model.rir_model._vq._embedding.weight, w_before)
    assert "rir_model._vq._embedding.weight" in model.state_dict()
    assert b200vq.swap_quantizers(model) == 0


def test_reference_state_dict_loads():
    from oracle.vq_oracle import import_reference_class
    import b200vq
    Ref = import_reference_class()
    if Ref is None:
        pytest.skip("/root/reference absent")
    ref = Ref(64, 16, 0.25)
    mine = b200vq.VectorQuantizer(64, 16, 0.25)
    mine.load_state_dict(ref.state_dict())
    assert torch.equal(mine._embedding.weight, ref._embedding.weight)


def test_whole_model_pickle_round_trip_then_swap(tmp_path):
    """The reference checkpoints are whole pickled modules (train_speech.py:117-118 `torch.save(model, path)`, reloaded with
    `torch.load(path)` at train_echoed_speech.py:18-19 / train_location.py:38): a model saved that way must load, swap its
    quantizers for the B200 ones and keep codebooks, frozen flags and state-dict keys -- and the swapped model must itself
    survive torch.save / torch.load (scratch buffers and process groups do not travel)."""
    from oracle.vq_oracle import import_reference_class
    import b200vq
    if import_reference_class() is None:
        pytest.skip("/root/reference absent")
    from acoustic_locating_vq_vae.vq_vae.convolutional_vq_vae import ConvolutionalVQVAE
    from acoustic_locating_vq_vae.vq_vae.echoed_speech_model import EchoedSpeechReconModel
    torch.manual_seed(21)
    rir = ConvolutionalVQVAE(20, 16, 8, 1, 8, 0.25, 32, use_jitter=False, out_channels=1)
    speech = ConvolutionalVQVAE(12, 16, 8, 1, 8, 0.25, 32)
    model = EchoedSpeechReconModel(rir, speech, 12, 16, 1, 8, True)
    path = tmp_path / "model_echoed_speech.pt"
    torch.save(model, path)                                             # exactly what the scripts do
    loaded = torch.load(path, weights_only=False)
    keys = sorted(loaded.state_dict().keys())
    codebooks = {n: p.detach().clone() for n, p in loaded.named_parameters() if n.endswith("_vq._embedding.weight")}
    assert len(codebooks) == 2
    assert b200vq.swap_quantizers(loaded) == 2
    assert sorted(loaded.state_dict().keys()) == keys                   # same checkpoint layout after the swap
    for n, w in codebooks.items():
        assert torch.equal(dict(loaded.named_parameters())[n], w)
    assert loaded.rir_model._vq._train_vq is False and loaded.speech_model._vq._train_vq is False
    path2 = tmp_path / "model_swapped.pt"
    torch.save(loaded, path2)
    again = torch.load(path2, weights_only=False)
    assert isinstance(again.rir_model._vq, b200vq.VectorQuantizer) and again.rir_model._vq._bufs is not None
    assert sorted(again.state_dict().keys()) == keys
    for n, w in codebooks.items():
        assert torch.equal(dict(again.named_parameters())[n], w)
    # and the reference class loads a state dict written by the swapped model (checkpoints stay exchangeable)
    ref_speech = ConvolutionalVQVAE(12, 16, 8, 1, 8, 0.25, 32)
    ref_speech.load_state_dict(again.speech_model.state_dict())
