"""Generate tests/golden/*.npz from the UNMODIFIED reference (build container only; /root/reference needed).

    python oracle/make_golden.py

The weights and inputs are rebuilt from seeds by novic_b200.synth (numpy PCG64, bit-reproducible anywhere), so
the fixtures only hold the reference's *outputs*: token ids, padding, scores, losses, and per-position logits
statistics (log-sum-exp, arg-max, top-2 margin, the target's logit and 64 probed vocabulary columns) instead of
the full B x C x V tensors.  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refload  # noqa: E402
from novic_b200 import synth  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
PROBE_SEED = 99
NUM_PROBES = 64
B_GOLD = 32


def weight_cases(dims):
    lively = synth.synth_state_dict(dims, seed=2, token_scale=0.25, jitter_norms=True)
    return {
        "lively": lively,
        "eos": synth.make_eos_friendly(lively, dims, beta=0.1),
        "eosall": synth.make_eos_ragged(lively),
    }


def probe_columns(V):
    return np.random.default_rng(PROBE_SEED).choice(V, size=NUM_PROBES, replace=False).astype(np.int64)


def logits_summary(logits: torch.Tensor, target: torch.Tensor, probes: np.ndarray) -> dict:
    top2 = logits.topk(2, dim=-1)
    return dict(
        lse=torch.logsumexp(logits, dim=-1).numpy(), argmax=top2.indices[..., 0].numpy(),
        top1=top2.values[..., 0].numpy(), margin=(top2.values[..., 0] - top2.values[..., 1]).numpy(),
        at_target=logits.gather(-1, target.clamp(min=0).unsqueeze(-1)).squeeze(-1).numpy(),
        probes=logits[..., torch.from_numpy(probes)].numpy(), mean=logits.mean(dim=-1).numpy())


def main():
    ref = refload.import_reference()
    dims = synth.DecoderDims()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    probes = probe_columns(dims.vocab_size)
    embed = synth.synth_embeddings(B_GOLD, seed=1234)
    out = {}
    for tag, sd in weight_cases(dims).items():
        model = refload.build_reference_decoder(ref, sd)
        with torch.inference_mode():
            # teacher-forced forward (embedding_decoder.py:659-777)
            tgt, pad = synth.synth_targets(B_GOLD, dims, seed=5)
            logits, epad, ls, lb, cor = model(embed, tgt, pad, None, True, True, False, None)
            for k, v in logits_summary(logits, tgt, probes).items():
                out[f"{tag}/tf/{k}"] = v
            out[f"{tag}/tf/loss"] = np.array([ls.item(), float(lb)], dtype=np.float64)
            out[f"{tag}/tf/correct"] = cor.numpy()
            out[f"{tag}/tf/effpad"] = epad.numpy()
            # guided correctness evaluation (embedding_decoder.py:754-760): targets drawn from a guide set, one of them corrupted
            ggt = synth.synth_guide_targets(300, dims, seed=21, first_pool=24)
            gidx = torch.randint(0, 300, (B_GOLD,), generator=torch.Generator().manual_seed(1))
            gtgt = ggt[gidx].clone()
            gtgt[5, 2] = 77
            gpad = torch.zeros_like(gtgt, dtype=torch.bool)
            gpad[:, 1:] = (gtgt[:, :-1] == 0).cummax(dim=1).values
            glog, _, _, _, gcor = model(embed, gtgt, gpad, None, True, True, False, ggt)
            out[f"{tag}/tfg/correct"] = gcor.numpy()
            gT = ggt.t()
            mism = torch.cat((torch.zeros(B_GOLD, 1, 300, dtype=torch.bool), (gtgt[:, :-1, None] != gT[None, :-1, :]).cummax(dim=1).values), dim=1)
            gs = torch.full((B_GOLD, gtgt.shape[1], dims.vocab_size + 1), float("-inf")).scatter_(2, gT[None].expand(B_GOLD, -1, -1).masked_fill(mism, dims.vocab_size), 0.0)[:, :, :-1]
            top2 = (gs + glog).topk(2, dim=-1).values
            out[f"{tag}/tfg/margin"] = (top2[..., 0] - top2[..., 1]).nan_to_num(nan=float("inf"), posinf=float("inf")).numpy()
            # multi-target weighted forward
            tgt3, pad3 = synth.synth_targets(8, dims, seed=6, multi=3)
            w3 = torch.from_numpy(np.random.default_rng(8).random((8, 3)).astype(np.float32))
            w3[1, 2] = 0.0
            model_m = refload.build_reference_decoder(ref, sd, multi_target=True, use_weights=True)
            lm, pm, lsm, lbm, corm = model_m(embed[:8], tgt3, pad3, w3, True, True, False, None)
            out[f"{tag}/tfm/lse"] = torch.logsumexp(lm, dim=-1).numpy()
            out[f"{tag}/tfm/probes"] = lm[..., torch.from_numpy(probes)].numpy()
            out[f"{tag}/tfm/loss"] = np.array([lsm.item(), lbm.item()], dtype=np.float64)
            out[f"{tag}/tfm/effpad"] = pm.numpy()
            out[f"{tag}/tfm/correct"] = corm.numpy()
            # greedy (embedding_decoder.py:779-850) exactly as infer.py:567-576 calls it, plus a tau/alpha variant
            for name, tau, alpha in (("g10", 1.0, 0.0), ("g07", 0.7, 0.5)):
                t, p, lg, gls, glb, sc = model.generate(embed, True, True, tau, alpha, None, None, False)
                out[f"{tag}/{name}/tok"] = t.numpy()
                out[f"{tag}/{name}/pad"] = p.numpy()
                out[f"{tag}/{name}/score"] = sc.numpy()
                out[f"{tag}/{name}/loss"] = np.array([gls.item(), float(glb)], dtype=np.float64)
                top2 = lg.topk(2, dim=-1).values
                out[f"{tag}/{name}/margin"] = (top2[..., 0] - top2[..., 1]).numpy()
                first = lg[:, 0, 1:].topk(2, dim=-1).values  # first step excludes the end token
                out[f"{tag}/{name}/margin"][:, 0] = (first[:, 0] - first[:, 1]).numpy()
                out[f"{tag}/{name}/lse"] = torch.logsumexp(lg, dim=-1).numpy()
                out[f"{tag}/{name}/probes"] = lg[..., torch.from_numpy(probes)].numpy()
            # beam (embedding_decoder.py:852-984)
            for name, H, tau, alpha in (("b3", 3, 1.0, 0.0), ("b5", 5, 1.3, 0.6), ("b10", 10, 1.0, 0.0)):
                t, p, sc = model.generate_beam(embed, H, tau, alpha, None, False, 0.0, None, False)
                out[f"{tag}/{name}/tok"] = t.numpy()
                out[f"{tag}/{name}/pad"] = p.numpy()
                out[f"{tag}/{name}/score"] = sc.numpy()
            # guided decoding (embedding_decoder.py:802-813, :915-920, :942-943): synthetic guide vocabularies, flat and deep tries
            for gname, W, pool in (("gflat", 300, 0), ("gdeep", 400, 24)):
                gt = synth.synth_guide_targets(W, dims, seed=21, first_pool=pool)
                for rname, renorm in (("p", False), ("r", True)):
                    t, p, lg, gls, glb, sc = model.generate(embed, True, True, 0.8, 0.3, None, gt, renorm)
                    out[f"{tag}/{gname}/greedy_{rname}/tok"] = t.numpy()
                    out[f"{tag}/{gname}/greedy_{rname}/pad"] = p.numpy()
                    out[f"{tag}/{gname}/greedy_{rname}/score"] = sc.numpy()
                    out[f"{tag}/{gname}/greedy_{rname}/loss"] = np.array([gls.item(), float(glb)], dtype=np.float64)
                    for H in (3, 10):   # beam_k10_vnone_gp_t1_a0 is the reference's default generation config (infer.py:55)
                        t, p, sc = model.generate_beam(embed, H, 1.0, 0.0, None, False, 0.0, gt, renorm)
                        out[f"{tag}/{gname}/beam{H}_{rname}/tok"] = t.numpy()
                        out[f"{tag}/{gname}/beam{H}_{rname}/pad"] = p.numpy()
                        out[f"{tag}/{gname}/beam{H}_{rname}/score"] = sc.numpy()
            # beam search with a vocabulary prior (embedding_decoder.py:881-891, :924-936): vocabulary alone, vocabulary = guide,
            # vocabulary != guide (shares 250 nouns with it); counts and per-token modes
            gt = synth.synth_guide_targets(400, dims, seed=21, first_pool=24)
            vt = torch.cat((gt[:250], synth.synth_guide_targets(200, dims, seed=25, first_pool=24)))
            for vname, H, g, v, per_token, scaler, renorm, tau, alpha in (
                    ("vonly_c", 3, None, vt, False, 0.6, False, 1.0, 0.0), ("vonly_t", 10, None, vt, True, 0.4, False, 0.9, 0.3),
                    ("vguide_c", 10, gt, gt, False, 0.5, False, 1.0, 0.0), ("vguide_t", 3, gt, gt.clone(), True, 0.8, True, 1.0, 0.0),
                    ("vdiff_c", 10, gt, vt, False, 0.5, True, 1.0, 0.0), ("vdiff_t", 3, gt, vt, True, 0.3, False, 1.2, 0.5)):
                t, p, sc = model.generate_beam(embed, H, tau, alpha, v, per_token, scaler, g, renorm)
                out[f"{tag}/{vname}/tok"] = t.numpy()
                out[f"{tag}/{vname}/pad"] = p.numpy()
                out[f"{tag}/{vname}/score"] = sc.numpy()
            # generate_all (embedding_decoder.py:986-1079): all guide targets scored by teacher forcing; K = W keeps the whole ranking
            if tag != "eosall":
                gt = synth.synth_guide_targets(200, dims, seed=23, first_pool=24)
                vt = torch.cat((gt[:150], synth.synth_guide_targets(120, dims, seed=24, first_pool=24)))
                for aname, renorm, vocab, per_token, scaler, tau, alpha in (
                        ("plain", False, None, False, 0.0, 1.0, 0.0), ("renorm", True, None, False, 0.0, 0.8, 0.4),
                        ("vcount", True, vt, False, 0.7, 1.0, 0.0), ("vtoken", False, vt, True, 0.5, 1.0, 0.3)):
                    t, p, sc = model.generate_all(embed[:16], gt.shape[0], tau, alpha, vocab, per_token, scaler, gt, renorm)
                    out[f"{tag}/all_{aname}/score"] = sc.numpy()
                    out[f"{tag}/all_{aname}/tok10"] = t[:, :10].numpy()
                    out[f"{tag}/all_{aname}/pad10"] = p[:, :10].numpy()
    # embedding noise (embedding_noise.py): outputs of the reference modules and the draws they consumed
    N = ref.embedding_noise
    e0 = synth.synth_embeddings(16, seed=9)
    F = e0.shape[1]
    def draws(seed, spec):
        torch.manual_seed(seed)
        res = []
        for kind in spec:
            res.append(torch.randn_like(e0) if kind == "N" else torch.randn(16, 1) if kind == "n" else torch.rand(16, 1))
        return res
    torch.manual_seed(11); out["noise/gauss_elem/out"] = N.GaussElemNoise(F, 3.25)(e0.clone()).numpy()
    out["noise/gauss_elem/na"] = draws(11, "N")[0].numpy()
    torch.manual_seed(12); out["noise/gauss_vec/out"] = N.GaussVecNoise(F, 0.8)(e0.clone()).numpy()
    na, g = draws(12, "Nn"); out["noise/gauss_vec/na"] = na.numpy(); out["noise/gauss_vec/ra"] = g.numpy().ravel()
    torch.manual_seed(13); out["noise/uniform_angle/out"] = N.UniformAngleNoise(F, 45.0, 75.0)(e0.clone()).numpy()
    na, u = draws(13, "Nu"); out["noise/uniform_angle/na"] = na.numpy(); out["noise/uniform_angle/ra"] = u.numpy().ravel()
    torch.manual_seed(14); out["noise/gauss_angle/out"] = N.GaussAngleNoise(F, 30.0, 40.0)(e0.clone()).numpy()
    na, g = draws(14, "Nn"); out["noise/gauss_angle/na"] = na.numpy(); out["noise/gauss_angle/ra"] = g.numpy().ravel()
    torch.manual_seed(15); out["noise/mix/out"] = N.GaussElemUniformAngleNoise(F, 3.25, 45.0, 75.0, 0.5)(e0.clone()).numpy()
    na, ua, ne, um = draws(15, "NuNu")
    out["noise/mix/na"] = na.numpy(); out["noise/mix/ra"] = ua.numpy().ravel(); out["noise/mix/nb"] = ne.numpy(); out["noise/mix/rb"] = um.numpy().ravel()
    out["meta/probes"] = probes
    path = os.path.join(GOLDEN_DIR, "reference_outputs.npz")
    np.savez_compressed(path, **{k.replace("/", "__"): v for k, v in out.items()})
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB, {len(out)} arrays")
    for tag in ("lively", "eos", "eosall"):
        p = out[f"{tag}/g10/pad"]
        print(tag, "greedy T =", p.shape[1], "rows finished early =", int(p.any(axis=1).sum()), "beam T =", out[f"{tag}/b3/tok"].shape[2])


if __name__ == "__main__":
    main()
