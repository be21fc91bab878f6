"""Loader for tests/golden/reference_outputs.npz (outputs of the unmodified reference, see oracle/make_golden.py)."""
import os

import numpy as np
import torch

from novic_b200 import synth

GOLDEN_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_outputs.npz")
B_GOLD = 32


class Golden:
    def __init__(self):
        self._z = np.load(GOLDEN_PATH)

    def __getitem__(self, key: str) -> torch.Tensor:
        return torch.from_numpy(self._z[key.replace("/", "__")])

    def has(self, key: str) -> bool:
        return key.replace("/", "__") in self._z.files


def weight_case(tag: str, dims: synth.DecoderDims = synth.DecoderDims()) -> dict:
    lively = synth.synth_state_dict(dims, seed=2, token_scale=0.25, jitter_norms=True)
    if tag == "lively":
        return lively
    if tag == "eos":
        return synth.make_eos_friendly(lively, dims, beta=0.1)
    if tag == "eosall":
        return synth.make_eos_ragged(lively)
    raise KeyError(tag)


def gold_embed() -> torch.Tensor:
    return synth.synth_embeddings(B_GOLD, seed=1234)


def guided_eval_case(dims: synth.DecoderDims = synth.DecoderDims()):
    """Guide set, targets drawn from it (one corrupted) and their padding - the inputs of the `tfg/*` fixtures (oracle/make_golden.py)."""
    gt = synth.synth_guide_targets(300, dims, seed=21, first_pool=24)
    idx = torch.randint(0, 300, (B_GOLD,), generator=torch.Generator().manual_seed(1))
    tgt = gt[idx].clone()
    tgt[5, 2] = 77
    pad = torch.zeros_like(tgt, dtype=torch.bool)
    pad[:, 1:] = (tgt[:, :-1] == 0).cummax(dim=1).values
    return gt, tgt, pad
