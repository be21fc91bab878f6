#!/usr/bin/env python
"""TEST INFRASTRUCTURE: golden vectors for tests/test_targets.py from the UNMODIFIED reference Embedder (embedders.py:169-254,
:331-406) driven by the toy tokenizer of that test.  Run in the build container (needs /root/reference):
    python oracle/make_targets_golden.py        -> tests/golden/targets_expected.npz
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refload  # noqa: E402
from tests import test_targets as tt  # noqa: E402  (toy tokenizer, noun list, switch grid)


def main():
    ref = refload.import_reference()
    out = {}
    for name in sorted(tt.TOKENIZERS):
        spec = tt.TOKENIZERS[name]
        emb = tt._toy_embedder(ref, spec)
        for si, sw in enumerate(tt.SWITCHES):
            cfg = emb.create_target_config(tt.NOUNS, **sw)
            emb.configure_target(cfg, tt.NOUNS)
            key = f"{name}__{si}"
            none = -(2 ** 31)
            out[f"{key}__scalars"] = np.array([cfg.vocab_size, none if cfg.start_token_id is None else cfg.start_token_id,
                                               none if cfg.end_token_id is None else cfg.end_token_id, cfg.pad_token_id, cfg.token_length], dtype=np.int64)
            if cfg.compact_ids:
                out[f"{key}__map"] = cfg.compact_map.numpy()
                out[f"{key}__unmap"] = cfg.compact_unmap.numpy()
            for bi, batch in enumerate(tt.BATCHES):
                ids, mask = emb.tokenize_target(batch)
                out[f"{key}__b{bi}__ids"] = ids.numpy()
                if mask is not None:
                    out[f"{key}__b{bi}__mask"] = mask.numpy()
                out[f"{key}__b{bi}__raw"] = emb.detokenize_target(ids).numpy()
    path = os.path.join(ROOT, "tests", "golden", "targets_expected.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
