"""Generate tests/golden/cache_*.bin / cache_expected.npz (build container only; /root/reference needed).

    python oracle/make_cache_golden.py

The cache files are written by novic_b200.cache.write_cache in the reference's documented format (embedding_cache.py:24-31); the
expected arrays are what the UNMODIFIED reference reader returns for them - EmbeddingCache.get_samples (embedding_cache.py:699-723)
and EmbeddingCache.Dataset.__getitem__ (:827-895) opened with strict_embedder=False (the files carry no embedder hash).
TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refload  # noqa: E402
from novic_b200 import cache  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
CASES = {"multi": dict(N=37, F=16, R1=9, C=6, M=3, weights=True, seed=0), "single": dict(N=21, F=8, R1=5, C=7, M=1, weights=False, seed=1)}


def synth_case(N, F, R1, C, M, weights, seed):
    """Small synthetic cache contents: ragged noun tokenisations, trailing unused target slots, descending non-negative weights."""
    g = torch.Generator().manual_seed(seed)
    emb = torch.nn.functional.normalize(torch.randn(N, F, generator=g), dim=-1)
    nouns = [f"noun {i}" for i in range(R1)]
    tok = torch.randint(1, 100, (R1, C), dtype=torch.int64, generator=g)
    lens = torch.randint(1, C - 1, (R1,), generator=g)
    mask = torch.arange(C).unsqueeze(0) > lens.unsqueeze(1)          # the end token at position `len` is not padding
    tok[mask] = 0
    tok[torch.arange(R1), lens] = 0
    et = torch.randint(1, R1 + 1, (N, M), dtype=torch.int32, generator=g)
    if M > 1:
        et[::3, M - 1] = 0
        et[::6, 1:] = 0
    w = None
    if weights:
        w = torch.rand(N, M, generator=g).sort(dim=1, descending=True).values
        w[et == 0] = 0
        w = w / w.sum(dim=1, keepdim=True)
    return emb, nouns, tok, mask, et, w


def stub_embedder(ref, C, F):
    tc = ref.embedders.TargetConfig(vocab_size=100, token_dtype=torch.int64, mask_dtype=torch.bool, start_token_id=None, end_token_id=0, pad_token_id=0,
                                    compact_ids=True, compact_map=None, compact_unmap=None, fixed_token_length=False, token_length=C, use_masks=True)
    return types.SimpleNamespace(target_config=tc, embed_dim=F, embed_dtype=torch.float32, token_dtype=torch.int64, vocab_size=100)


def reference_outputs(ref, path, C, F, batch_size):
    import embedding_cache
    out = {}
    ec = embedding_cache.EmbeddingCache(path, stub_embedder(ref, C, F), use_targets=None, strict_embedder=False)
    with ec:
        for name, (a, b) in {"head": (0, 5), "mid": (5, 12), "tail": (len(ec) - 3, len(ec) + 4), "empty": (7, 7)}.items():
            for k, t in zip(("embed", "ids", "target", "mask", "weight"), ec.get_samples(a, b)):
                out[f"samples/{name}/{k}"] = t.clone().numpy()
        for mode, training, offset in (("eval", False, 0), ("train0", True, 0), ("train5", True, 5), ("trainwrap", True, len(ec) - 3)):
            ds = embedding_cache.EmbeddingCache.Dataset(ec, batch_size=batch_size, training=training)
            ds.data_config = ds.nominal_data_config
            ds.epoch_index_offset = offset
            out[f"batches/{mode}/count"] = np.array([len(ds)])
            for i in range(len(ds)):
                for k, t in zip(("embed", "target", "mask", "weight"), ds[i]):
                    if t is not None:
                        out[f"batches/{mode}/{i}/{k}"] = t.clone().numpy()
    return out


def main():
    ref = refload.import_reference()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    allout = {}
    for name, cfg in CASES.items():
        emb, nouns, tok, mask, et, w = synth_case(**cfg)
        path = os.path.join(GOLDEN_DIR, f"cache_{name}.bin")
        cache.write_cache(path, emb, nouns, tok, mask, et, w)
        for k, v in reference_outputs(ref, path, cfg["C"], cfg["F"], batch_size=8).items():
            allout[f"{name}/{k}"] = v
        print(name, os.path.getsize(path), "bytes")
    np.savez_compressed(os.path.join(GOLDEN_DIR, "cache_expected.npz"), **{k.replace("/", "__"): v for k, v in allout.items()})
    print("arrays:", len(allout))


if __name__ == "__main__":
    main()
