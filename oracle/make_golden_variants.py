"""Generate tests/golden/reference_variants.npz from the UNMODIFIED reference (build container only; /root/reference needed).

    python oracle/make_golden_variants.py

Where oracle/make_golden.py pins the default configuration (F = 1024, V = 6912, Cmax = 16), this file pins the constructor
values a released checkpoint or a training configuration can carry besides (embedding_decoder.py:43-75, :633-640):
embedding sizes 768 / 1152 (the SigLIP B/16 and SO400M checkpoints, README.md:293-300), a vocabulary that is not a multiple
of the 64 / 128-column logits tiles (with and without `vocab_quant`), `label_smoothing`, `num_end_loss = 2`,
`strictly_causal`, and a shorter `token_length`.  It also stores
  - gradients of the reference (`model.train()`, dropout 0, `loss_sum.backward()`) as per-tensor norms, probed elements and
    random projections (the full 12.7 M-element gradients would be 51 MB per case),
  - the first 256 rows of the BASELINE config #2 batch (seed-1 weights, 4096 seed-1234 embeddings) decoded by the reference,
  - per-sample pruning margins of the beam searches of reference_outputs.npz (from the oracle restatement, after checking
    that it reproduces the reference's beams bit for bit on those inputs).
Weights and inputs are rebuilt from seeds by novic_b200.synth; only outputs are stored.  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import novic_oracle as orc  # noqa: E402
from oracle import refload  # noqa: E402
from novic_b200 import synth  # noqa: E402

from tests.golden_util import (B_GOLD, BASE_DIMS as BASE, GRAD_CASES, VARIANTS, grad_case_inputs, grad_probe, probe_columns,  # noqa: E402
                               variant_state_dict)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def grad_summary(named_grads: dict) -> dict:
    out = {}
    for i, k in enumerate(sorted(named_grads)):
        g = named_grads[k].detach().double().flatten()
        idx, proj = grad_probe(i, g.numel())
        out[f"{k}/norm"] = np.array(g.norm().item())
        out[f"{k}/probe"] = g[torch.from_numpy(idx)].numpy()
        out[f"{k}/proj"] = (torch.from_numpy(proj).double() @ g).numpy()
    return out


def build(ref, dims, overrides, sd, **kw):
    return refload.build_reference_decoder(ref, sd, vocab_size=dims.vocab_size, token_length=dims.token_length,
                                           embed_dim=dims.embed_dim, **kw, **overrides)


def main():
    ref = refload.import_reference()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    out = {}
    for name, (dims, overrides) in VARIANTS.items():
        sd = variant_state_dict(dims, overrides)
        model = build(ref, dims, overrides, sd)
        probes = probe_columns(dims.vocab_size)
        pt = torch.from_numpy(probes)
        embed = synth.synth_embeddings(B_GOLD, dims.embed_dim, seed=1234)
        with torch.inference_mode():
            tgt, pad = synth.synth_targets(B_GOLD, dims, seed=5)
            logits, epad, ls, lb, cor = model(embed, tgt, pad, None, True, True, False, None)
            assert logits.shape[-1] == dims.vocab_size
            top2 = logits.topk(2, dim=-1)
            out[f"{name}/tf/lse"] = torch.logsumexp(logits, -1).numpy()
            out[f"{name}/tf/argmax"] = top2.indices[..., 0].numpy()
            out[f"{name}/tf/margin"] = (top2.values[..., 0] - top2.values[..., 1]).numpy()
            out[f"{name}/tf/at_target"] = logits.gather(-1, tgt.unsqueeze(-1)).squeeze(-1).numpy()
            out[f"{name}/tf/probes"] = logits[..., pt].numpy()
            out[f"{name}/tf/last_cols"] = logits[..., -8:].numpy()          # the ragged end of the vocabulary
            out[f"{name}/tf/loss"] = np.array([ls.item(), float(lb)], dtype=np.float64)
            out[f"{name}/tf/correct"] = cor.numpy()
            out[f"{name}/tf/effpad"] = epad.numpy()
            t, p, lg, gls, glb, sc = model.generate(embed, True, True, 1.0, 0.0, None, None, False)
            out[f"{name}/g10/tok"] = t.numpy()
            out[f"{name}/g10/pad"] = p.numpy()
            out[f"{name}/g10/score"] = sc.numpy()
            out[f"{name}/g10/loss"] = np.array([gls.item(), float(glb)], dtype=np.float64)
            t2 = lg.topk(2, dim=-1).values
            out[f"{name}/g10/margin"] = (t2[..., 0] - t2[..., 1]).numpy()
            first = lg[:, 0, 1:].topk(2, dim=-1).values
            out[f"{name}/g10/margin"][:, 0] = (first[:, 0] - first[:, 1]).numpy()
            out[f"{name}/g10/lse"] = torch.logsumexp(lg, -1).numpy()
            out[f"{name}/g10/probes"] = lg[..., pt].numpy()
            if name != "nel2":   # num_end_loss > 1 lets a finished beam grow one more (discarded) token: refused by the product
                t, p, sc = model.generate_beam(embed, 3, 1.0, 0.0, None, False, 0.0, None, False)
                out[f"{name}/b3/tok"] = t.numpy()
                out[f"{name}/b3/pad"] = p.numpy()
                out[f"{name}/b3/score"] = sc.numpy()
        print(name, "greedy T =", out[f"{name}/g10/tok"].shape[1], "finished rows =", int(out[f"{name}/g10/pad"].any(axis=1).sum()),
              "loss", out[f"{name}/tf/loss"])
    # gradients of the unmodified reference: training mode, dropout 0 (torch's dropout stream is not reproducible elsewhere)
    for name, (dims, overrides, multi) in GRAD_CASES.items():
        sd = variant_state_dict(dims, overrides)
        kw = dict(multi_target=True, use_weights=True) if multi else {}
        model = build(ref, dims, {**overrides, "input_dropout": 0.0, "layer_dropout": 0.0}, sd, **kw).train()
        embed, tgt, pad, w = grad_case_inputs(dims, multi)
        _, _, ls, lb, cor = model(embed, tgt, pad, w, True, True, False, None)
        ls.backward()
        grads = {k: p.grad for k, p in model.named_parameters()}
        assert all(g is not None for g in grads.values())
        out[f"{name}/loss"] = np.array([ls.item(), float(lb)], dtype=np.float64)
        out[f"{name}/correct"] = cor.numpy()
        for k, v in grad_summary(grads).items():
            out[f"{name}/{k}"] = v
        if overrides.get("vocab_quant"):
            assert float(grads["logits_linear.weight"][dims.vocab_size:].abs().max()) == 0.0
        print(name, "loss", out[f"{name}/loss"], "|g| tied", float(out[f"{name}/logits_linear.weight/norm"]))
    # BASELINE config #2: the bench batch itself (first 256 of the 4096 embeddings), decoded by the reference
    sd1 = synth.synth_state_dict(BASE, seed=1)
    model = build(ref, BASE, {}, sd1)
    e = synth.synth_embeddings(4096, seed=1234)[:256]
    with torch.inference_mode():
        t, p, lg, gls, glb, sc = model.generate(e, True, True, 1.0, 0.0, None, None, False)
    out["bench256/tok"] = t.numpy()
    out["bench256/pad"] = p.numpy()
    out["bench256/score"] = sc.numpy()
    t2 = lg.topk(2, dim=-1).values
    out["bench256/margin"] = (t2[..., 0] - t2[..., 1]).numpy()
    first = lg[:, 0, 1:].topk(2, dim=-1).values
    out["bench256/margin"][:, 0] = (first[:, 0] - first[:, 1]).numpy()
    out["bench256/lse"] = torch.logsumexp(lg, -1).numpy()
    print("bench256: T =", t.shape[1], "distinct first tokens =", len(np.unique(t[:, 0].numpy())))
    # pruning margins of the main fixtures' beam searches
    main_gold = np.load(os.path.join(GOLDEN_DIR, "reference_outputs.npz"))
    lively = synth.synth_state_dict(BASE, seed=2, token_scale=0.25, jitter_norms=True)
    cases = {"lively": lively, "eos": synth.make_eos_friendly(lively, BASE, beta=0.1), "eosall": synth.make_eos_ragged(lively)}
    embed = synth.synth_embeddings(B_GOLD, seed=1234)
    for tag, sd in cases.items():
        cfg = orc.cfg_from_state_dict(sd)
        for bname, H, tau, alpha in (("b3", 3, 1.0, 0.0), ("b5", 5, 1.3, 0.6), ("b10", 10, 1.0, 0.0)):
            with torch.inference_mode():
                o = orc.generate_beam(cfg, sd, embed, H, tau, alpha)
            same = np.array_equal(o["target"].numpy(), main_gold[f"{tag}__{bname}__tok"]) and np.array_equal(o["padding"].numpy(), main_gold[f"{tag}__{bname}__pad"])
            assert same, f"oracle beams differ from the reference's for {tag}/{bname}: margins would be meaningless"
            out[f"{tag}/{bname}/prune_margin"] = o["margin"].numpy()
            print(tag, bname, "margin quantiles", np.quantile(o["margin"].numpy(), [0.1, 0.5, 0.9]).round(4))
    path = os.path.join(GOLDEN_DIR, "reference_variants.npz")
    np.savez_compressed(path, **{k.replace("/", "__"): v for k, v in out.items()})
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB, {len(out)} arrays")


if __name__ == "__main__":
    main()
