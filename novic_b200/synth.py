"""Synthetic weights and inputs for the decoder hot path (no checkpoints or datasets exist offline).

`synth_state_dict` draws a state dict with the reference's parameter names/shapes and
the standard deviations its default 'balanced' initialisation produces (embedding_decoder.py:203-226,
:228-278, :329-407).  The draws come from numpy's PCG64 stream so that the *same* weights can be rebuilt
bit-for-bit on any machine from (seed, shape) alone: the golden vectors under tests/golden/ were produced
by loading exactly these weights into the reference's own PrefixedIterDecoder.

`synth_embeddings` follows the recipe the reference uses for its random embedding cache
(embedding_cache_writers.py:43): unit-normalised standard normal vectors.
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np
import torch


@dataclasses.dataclass(frozen=True)
class DecoderDims:
    """Default architecture of config/train.yaml:224-308 at the BASELINE configs (F=1024, V=6912, Cmax=16)."""
    embed_dim: int = 1024
    hidden_dim: int = 512
    ffn_dim: int = 128
    num_layers: int = 6
    num_heads: int = 8
    prefix_len: int = 4
    vocab_size: int = 6912
    token_length: int = 16

    @property
    def max_seq_len(self) -> int:
        return self.prefix_len + self.token_length - 1


def init_stds(d: DecoderDims) -> dict:
    E, L, K, P = d.hidden_dim, d.num_layers, d.ffn_dim, d.prefix_len
    f = 1.0 / math.sqrt(E)
    per_layer = 1.0 / math.sqrt(2 * L)                 # init_tfrm_proj_layers
    attn_scale = math.sqrt((1 + (P - 1) / P) / P)      # self_attn_dim = P, nominal_std = 1
    gelu_gain = 0.6521                                 # utils.py:107
    return {
        "embed_mlp": 1.0 / math.sqrt(2.0),             # balanced init, no output bias -> 1/sqrt(2)
        "logits": 1.0 / math.sqrt(2.0),
        "pos": 1.0 / math.sqrt(2.0),
        "in_proj": f,
        "out_proj": f / attn_scale * per_layer,
        "linear1": f,
        "linear2": per_layer / (math.sqrt(K) * gelu_gain),
        "norm": 1.0,
        "final_norm": f,                               # init_tfrm_unit_postnorm
    }


def synth_state_dict(dims: DecoderDims = DecoderDims(), seed: int = 1, *, token_scale: float = 1.0,
                     jitter_norms: bool = False, dtype: torch.dtype = torch.float32) -> dict:
    """Random-init state dict, reproducible from `seed`.

    token_scale < 1 shrinks the tied token/logits matrix (a freshly initialised tied model otherwise repeats
    one token per row, which makes a poor parity test); jitter_norms perturbs the LayerNorm gains away from 1
    so that a kernel that forgot to apply them is caught.
    """
    rng = np.random.default_rng(seed)
    s = init_stds(dims)
    E, K, P, F, V = dims.hidden_dim, dims.ffn_dim, dims.prefix_len, dims.embed_dim, dims.vocab_size

    def normal(shape, std):
        return torch.from_numpy((rng.standard_normal(shape) * std).astype(np.float32)).to(dtype)

    def gain(n, base):
        if not jitter_norms:
            return torch.full((n,), base, dtype=dtype)
        return torch.from_numpy((base * (1.0 + 0.25 * rng.standard_normal(n))).astype(np.float32)).to(dtype)

    sd = {
        "embed_mlp.mlp.0.weight": normal((P * E, F), s["embed_mlp"]),
        "logits_linear.weight": normal((V, E), s["logits"] * token_scale),
        "pos_embedding.embedding.weight": normal((dims.max_seq_len, E), s["pos"]),
    }
    for l in range(dims.num_layers):
        p = f"transformer.layers.{l}."
        sd[p + "self_attn.in_proj_weight"] = normal((3 * E, E), s["in_proj"])
        sd[p + "self_attn.out_proj.weight"] = normal((E, E), s["out_proj"])
        sd[p + "linear1.weight"] = normal((K, E), s["linear1"])
        sd[p + "linear2.weight"] = normal((E, K), s["linear2"])
        sd[p + "norm1.weight"] = gain(E, s["norm"])
        sd[p + "norm2.weight"] = gain(E, s["norm"])
    sd["transformer.norm.weight"] = gain(E, s["final_norm"])
    S = dims.max_seq_len
    mask = torch.triu(torch.full((S, S), float("-inf"), dtype=dtype), diagonal=1)
    mask[:P, :P] = 0.0
    sd["causality_mask"] = mask
    return sd


def make_eos_friendly(sd: dict, dims: DecoderDims = DecoderDims(), beta: float = 0.8) -> dict:
    """Make the end token (id 0) reachable so that early-exit / padding branches are exercised:
    row 0 of the tied matrix becomes beta * sum of positional embeddings 6..13."""
    sd = dict(sd)
    w = sd["logits_linear.weight"].clone()
    w[0] = beta * sd["pos_embedding.embedding.weight"][6:14].sum(dim=0)
    sd["logits_linear.weight"] = w
    return sd


def make_eos_ragged(sd: dict, coeffs=(0.08, 0.12, 0.5)) -> dict:
    """End-token row built from three position bands: every sample finishes, at different steps, so that a greedy /
    beam decode takes the all-finished early exit (embedding_decoder.py:817-820, :964-967) with ragged lengths."""
    sd = dict(sd)
    pos = sd["pos_embedding.embedding.weight"]
    w = sd["logits_linear.weight"].clone()
    w[0] = coeffs[0] * pos[6:9].sum(dim=0) + coeffs[1] * pos[9:12].sum(dim=0) + coeffs[2] * pos[12:16].sum(dim=0)
    sd["logits_linear.weight"] = w
    return sd


def synth_embeddings(batch: int, embed_dim: int = 1024, seed: int = 1234, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((batch, embed_dim)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return torch.from_numpy(x).to(dtype)


def synth_targets(batch: int, dims: DecoderDims = DecoderDims(), seed: int = 7, multi: int = 0):
    """Random training targets: ids in [1, V), length ~ U[2, Cmax] including the end token 0, then padding
    (SURVEY §8d config #4).  Returns (target int64, padding bool) shaped B x C, or B x M x C when multi > 0."""
    rng = np.random.default_rng(seed)
    n = batch * max(multi, 1)
    C = dims.token_length
    tok = rng.integers(1, dims.vocab_size, size=(n, C), dtype=np.int64)
    length = rng.integers(2, C + 1, size=(n,))
    col = np.arange(C)[None, :]
    tok[col >= (length[:, None] - 1)] = 0
    pad = col >= length[:, None]
    t, p = torch.from_numpy(tok), torch.from_numpy(pad)
    if multi > 0:
        t, p = t.view(batch, multi, C), p.view(batch, multi, C)
    return t, p


def synth_guide_targets(num_targets: int, dims: DecoderDims = DecoderDims(), seed: int = 21, max_tokens: int = 6,
                        first_pool: int = 0) -> torch.Tensor:
    """Synthetic guide vocabulary (infer.py:687-710 builds the real one by tokenising noun strings): W unique rows of
    1..max_tokens token ids in [1, V) followed by the end token 0 and zero padding, shaped W x Cmax int64.
    first_pool > 0 draws the tokens from a pool of that many ids so that the nouns share prefixes (a deep trie)."""
    rng = np.random.default_rng(seed)
    C, V = dims.token_length, dims.vocab_size
    pool = rng.choice(np.arange(1, V), size=first_pool, replace=False) if first_pool > 0 else None
    rows = set()
    while len(rows) < num_targets:
        n = int(rng.integers(1, max_tokens + 1))
        ids = rng.choice(pool, size=n) if pool is not None else rng.integers(1, V, size=n)
        rows.add(tuple(int(t) for t in ids))
    out = np.zeros((num_targets, C), dtype=np.int64)
    for i, r in enumerate(sorted(rows)):
        out[i, :len(r)] = r
    return torch.from_numpy(out[rng.permutation(num_targets)])
