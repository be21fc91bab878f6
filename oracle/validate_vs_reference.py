"""Pin the oracle: run oracle/novic_oracle.py against the reference's own classes (build container only).

Usage: python oracle/validate_vs_reference.py            (prints one line per check, exits non-zero on failure)
TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import novic_oracle as orc  # noqa: E402
from oracle import refload  # noqa: E402
from novic_b200 import synth  # noqa: E402


def check(name, ok, detail=""):
    print(f"[{'ok' if ok else 'FAIL'}] {name} {detail}")
    return bool(ok)


def run() -> bool:
    torch.manual_seed(0)
    ref = refload.import_reference()
    dims = synth.DecoderDims()
    ok = True
    for tag, sd in (
        ("init", synth.synth_state_dict(dims, seed=1)),
        ("lively", synth.synth_state_dict(dims, seed=2, token_scale=0.25, jitter_norms=True)),
        ("eos", synth.make_eos_friendly(synth.synth_state_dict(dims, seed=1), dims, beta=0.8)),
    ):
        cfg = orc.cfg_from_state_dict(sd)
        model = refload.build_reference_decoder(ref, sd)
        embed = synth.synth_embeddings(24, seed=1234)
        with torch.inference_mode():
            # teacher-forced forward, random targets + padding
            tgt, pad = synth.synth_targets(24, dims, seed=5)
            r_logits, r_pad, r_ls, r_lb, r_cor = model(embed, tgt, pad, None, True, True, False, None)
            o_logits, o_ls, o_lb, o_cor = orc.forward_loss(cfg, sd, embed, tgt, pad, None)
            valid = ~pad
            err = (r_logits - o_logits)[valid].abs().max().item()
            ok &= check(f"{tag}: teacher-forced logits", err < 2e-4, f"max|d|={err:.2e}")
            ok &= check(f"{tag}: loss", abs(r_ls.item() - o_ls.item()) < 1e-3 * abs(r_ls.item()) and int(r_lb) == int(o_lb),
                        f"{r_ls.item():.4f} vs {o_ls.item():.4f} basis {int(r_lb)}")
            ok &= check(f"{tag}: correct", torch.equal(r_cor, o_cor))
            # multi-target weighted forward (B x M x C)
            tgt3, pad3 = synth.synth_targets(8, dims, seed=6, multi=3)
            w3 = torch.rand(8, 3)
            w3[1, 2] = 0.0
            model_m = refload.build_reference_decoder(ref, sd, multi_target=True, use_weights=True)
            rm = model_m(embed[:8], tgt3, pad3, w3, True, True, False, None)
            om = orc.forward_loss(cfg, sd, embed[:8], tgt3.view(24, -1), pad3.view(24, -1), w3.view(-1))
            effv = ~rm[1].view(24, -1)
            err = (rm[0].view(24, -1, cfg.vocab_size) - om[0])[effv].abs().max().item()
            ok &= check(f"{tag}: multi-target weighted", err < 2e-4 and abs(rm[2].item() - om[1].item()) < 1e-3 * abs(om[1].item())
                        and abs(rm[3].item() - om[2].item()) < 1e-4 * abs(om[2].item()),
                        f"max|d|={err:.2e} loss {rm[2].item():.4f}/{om[1].item():.4f} basis {rm[3].item():.4f}/{om[2].item():.4f}")
            # greedy
            for tau, alpha in ((1.0, 0.0), (0.7, 0.5)):
                r = model.generate(embed, True, True, tau, alpha, None, None, False)
                o = orc.generate_greedy(cfg, sd, embed, tau, alpha)
                same = torch.equal(r[0], o["target"]) and torch.equal(r[1], o["padding"])
                serr = (r[5] - o["score"]).abs().max().item()
                ok &= check(f"{tag}: greedy tau={tau} alpha={alpha}", same and serr < 2e-3 and abs(r[3].item() - o["loss_sum"].item()) < 2e-3 * abs(r[3].item())
                            and int(r[4]) == int(o["loss_basis"]),
                            f"T={r[0].shape[1]} score|d|={serr:.2e} finished={int(o['padding'].any(dim=1).sum())}")
            # beam
            for H, tau, alpha in ((3, 1.0, 0.0), (5, 1.3, 0.6)):
                r = model.generate_beam(embed, H, tau, alpha, None, False, 0.0, None, False)
                o = orc.generate_beam(cfg, sd, embed, H, tau, alpha)
                same = torch.equal(r[0], o["target"]) and torch.equal(r[1], o["padding"])
                serr = (r[2] - o["score"]).abs().max().item()
                ok &= check(f"{tag}: beam H={H} tau={tau} alpha={alpha}", same and serr < 2e-3, f"T={r[0].shape[2]} score|d|={serr:.2e}")
    # noise: feed identical draws by re-seeding torch's generator the way each reference module consumes it
    e0 = synth.synth_embeddings(32, seed=9)
    F = e0.shape[1]
    N = ref.embedding_noise
    torch.manual_seed(11); r = N.GaussElemNoise(F, 3.25)(e0.clone())
    torch.manual_seed(11); n = torch.randn_like(e0)
    ok &= check("noise GaussElem", (r - orc.noise_gauss_elem(e0, n, 3.25)).abs().max().item() < 1e-6)
    torch.manual_seed(12); r = N.GaussVecNoise(F, 0.8)(e0.clone())
    torch.manual_seed(12); n = torch.randn_like(e0); g = torch.randn(32, 1)
    ok &= check("noise GaussVec", (r - orc.noise_gauss_vec(e0, n, g, 0.8)).abs().max().item() < 1e-6)
    torch.manual_seed(13); r = N.UniformAngleNoise(F, 45.0, 75.0)(e0.clone())
    torch.manual_seed(13); n = torch.randn_like(e0); u = torch.rand(32, 1)
    ok &= check("noise UniformAngle", (r - orc.noise_angle(e0, n, orc.uniform_angle(u, 45.0, 75.0))).abs().max().item() < 1e-6)
    torch.manual_seed(14); r = N.GaussAngleNoise(F, 30.0, 40.0)(e0.clone())
    torch.manual_seed(14); n = torch.randn_like(e0); g = torch.randn(32, 1)
    ok &= check("noise GaussAngle", (r - orc.noise_angle(e0, n, orc.gauss_angle(g, 30.0, 40.0))).abs().max().item() < 1e-6)
    torch.manual_seed(15); r = N.GaussElemUniformAngleNoise(F, 3.25, 45.0, 75.0, 0.5)(e0.clone())
    torch.manual_seed(15); na = torch.randn_like(e0); ua = torch.rand(32, 1); ne = torch.randn_like(e0); um = torch.rand(32, 1)
    o = orc.noise_gauss_elem_uniform_angle(e0, na, ua, ne, um, 3.25, 45.0, 75.0, 0.5)
    ok &= check("noise GaussElemUniformAngle", (r - o).abs().max().item() < 1e-6)
    # init statistics of the product's synthetic weights vs a reference-constructed model
    torch.manual_seed(1)
    fresh = refload.build_reference_decoder(ref, None).state_dict()
    mine = synth.synth_state_dict(dims, seed=1)
    worst = 0.0
    for k, v in fresh.items():
        assert mine[k].shape == v.shape, k
        if v.numel() > 600 and k != "causality_mask":
            worst = max(worst, abs(mine[k].std().item() / v.std().item() - 1.0))
    ok &= check("synthetic init stds match reference init", worst < 0.03 and set(fresh) == set(mine), f"worst rel dev {worst:.3f}")
    ok &= check("causality_mask identical", torch.equal(fresh["causality_mask"], mine["causality_mask"]))
    for k in ("transformer.layers.0.norm1.weight", "transformer.norm.weight"):
        ok &= check(f"{k} identical", torch.allclose(fresh[k], mine[k]))
    return ok


if __name__ == "__main__":
    sys.exit(0 if run() else 1)
