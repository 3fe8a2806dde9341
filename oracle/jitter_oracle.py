"""CPU oracle for Jitter -- TEST INFRASTRUCTURE ONLY.

Restates /root/reference/src/acoustic_locating_vq_vae/vq_vae/modules/jitter.py:47-70.  Two things a port must
not "fix":
  * `replace = [True, False][np.random.choice([1, 0], p=[p, 1 - p])]` (:50): choice returns 1 with probability p,
    and [True, False][1] is False -- a column is REPLACED with probability 1 - p (0.75 at the decoder's p = 0.25);
  * the copy sources come from a clone taken before the loop (:48), so replacements never chain, and the input
    tensor is mutated IN PLACE and returned (:68-70).
`source_columns` reproduces the reference's np.random consumption call for call, so a given numpy seed yields the
same decisions.
"""
from __future__ import annotations

import numpy as np
import torch


def source_columns(length: int, probability: float) -> np.ndarray:
    """src[i] = the column of the ORIGINAL tensor that ends up in column i (jitter.py:50-67)."""
    src = np.arange(length, dtype=np.int64)
    for i in range(length):
        replace = [True, False][np.random.choice([1, 0], p=[probability, 1 - probability])]   # :50
        if replace:
            if i == 0:
                neighbor = i + 1                                                              # :52-53
            elif i == length - 1:
                neighbor = i - 1                                                              # :54-55
            else:
                neighbor = i + np.random.choice([-1, 1], p=[0.5, 0.5])                        # :62
            src[i] = neighbor
    return src


def jitter(quantized: torch.Tensor, probability: float) -> torch.Tensor:
    """In-place jitter of a (B, D, T) tensor; returns the same tensor object (jitter.py:47-70)."""
    original = quantized.detach().clone()
    src = source_columns(original.size(2), probability)
    for i, j in enumerate(src):
        if j != i:
            quantized[:, :, i] = original[:, :, j]
    return quantized


def check_against_reference(seed: int = 11, shape=(3, 5, 40), probability: float = 0.25) -> bool:
    import os, sys
    if not os.path.isdir("/root/reference"):
        return False
    for p in ("/root/reference", "/root/reference/src"):
        if p not in sys.path:
            sys.path.insert(0, p)
    from acoustic_locating_vq_vae.vq_vae.modules.jitter import Jitter      # type: ignore
    torch.manual_seed(seed)
    x = torch.randn(*shape)
    np.random.seed(seed)
    ref = Jitter(probability)(x.clone())
    np.random.seed(seed)
    mine = jitter(x.clone(), probability)
    assert torch.equal(ref, mine)
    return True
