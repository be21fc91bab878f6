"""The CPU oracle must reproduce the committed outputs of the unmodified reference (tests/golden/)."""
import numpy as np
import pytest
import torch

from novic_b200 import synth
from oracle import novic_oracle as orc
from tests.golden_util import guided_eval_case, B_GOLD, Golden, gold_embed, weight_case

TAGS = ("lively", "eos", "eosall")


@pytest.fixture(scope="module")
def gold():
    return Golden()


@pytest.mark.parametrize("tag", TAGS)
def test_teacher_forced_forward(gold, tag):
    dims = synth.DecoderDims()
    sd = weight_case(tag)
    cfg = orc.cfg_from_state_dict(sd)
    tgt, pad = synth.synth_targets(B_GOLD, dims, seed=5)
    with torch.inference_mode():
        logits, ls, lb, cor = orc.forward_loss(cfg, sd, gold_embed(), tgt, pad, None)
    valid = ~pad
    probes = gold["meta/probes"]
    assert (logits[..., probes] - gold[f"{tag}/tf/probes"])[valid].abs().max() < 2e-4
    assert (torch.logsumexp(logits, -1) - gold[f"{tag}/tf/lse"])[valid].abs().max() < 2e-4
    assert torch.equal(logits.argmax(-1)[valid], gold[f"{tag}/tf/argmax"][valid])
    assert torch.equal(cor, gold[f"{tag}/tf/correct"])
    assert abs(ls.item() - gold[f"{tag}/tf/loss"][0].item()) < 1e-3 * abs(ls.item())
    assert int(lb) == int(gold[f"{tag}/tf/loss"][1].item())


@pytest.mark.parametrize("tag", ("lively", "eos"))
def test_multi_target_weighted_forward(gold, tag):
    dims = synth.DecoderDims()
    sd = weight_case(tag)
    cfg = orc.cfg_from_state_dict(sd)
    tgt3, pad3 = synth.synth_targets(8, dims, seed=6, multi=3)
    w3 = torch.from_numpy(np.random.default_rng(8).random((8, 3)).astype(np.float32))
    w3[1, 2] = 0.0
    with torch.inference_mode():
        logits, ls, lb, cor = orc.forward_loss(cfg, sd, gold_embed()[:8], tgt3.view(24, -1), pad3.view(24, -1), w3.view(-1))
    effpad = gold[f"{tag}/tfm/effpad"].view(24, -1)
    valid = ~effpad
    assert (logits[..., gold["meta/probes"]] - gold[f"{tag}/tfm/probes"].view(24, 16, -1))[valid].abs().max() < 2e-4
    assert abs(ls.item() - gold[f"{tag}/tfm/loss"][0].item()) < 1e-3 * abs(ls.item())
    assert abs(lb.item() - gold[f"{tag}/tfm/loss"][1].item()) < 1e-4 * abs(lb.item())
    assert torch.equal(cor.view(8, 3, -1), gold[f"{tag}/tfm/correct"])


@pytest.mark.parametrize("tag", TAGS)
@pytest.mark.parametrize("name,tau,alpha", (("g10", 1.0, 0.0), ("g07", 0.7, 0.5)))
def test_greedy(gold, tag, name, tau, alpha):
    sd = weight_case(tag)
    cfg = orc.cfg_from_state_dict(sd)
    with torch.inference_mode():
        o = orc.generate_greedy(cfg, sd, gold_embed(), tau, alpha)
    assert torch.equal(o["target"], gold[f"{tag}/{name}/tok"])
    assert torch.equal(o["padding"], gold[f"{tag}/{name}/pad"])
    assert (o["score"] - gold[f"{tag}/{name}/score"]).abs().max() < 2e-3
    assert abs(o["loss_sum"].item() - gold[f"{tag}/{name}/loss"][0].item()) < 1e-3 * abs(o["loss_sum"].item())
    assert int(o["loss_basis"]) == int(gold[f"{tag}/{name}/loss"][1].item())


@pytest.mark.parametrize("tag", TAGS)
@pytest.mark.parametrize("name,H,tau,alpha", (("b3", 3, 1.0, 0.0), ("b5", 5, 1.3, 0.6), ("b10", 10, 1.0, 0.0)))
def test_beam(gold, tag, name, H, tau, alpha):
    sd = weight_case(tag)
    cfg = orc.cfg_from_state_dict(sd)
    with torch.inference_mode():
        o = orc.generate_beam(cfg, sd, gold_embed(), H, tau, alpha)
    assert torch.equal(o["target"], gold[f"{tag}/{name}/tok"])
    assert torch.equal(o["padding"], gold[f"{tag}/{name}/pad"])
    assert (o["score"] - gold[f"{tag}/{name}/score"]).abs().max() < 2e-3


def test_early_exit_fixture_is_exercised(gold):
    # the EOS-friendly weights must actually trigger the all-finished early exit (T < G) and ragged lengths
    assert gold["lively/g10/tok"].shape[1] == 15 and not gold["lively/g10/pad"].any()
    assert gold["eos/g10/tok"].shape[1] == 15                      # some samples never finish ...
    n_eos = (~gold["eos/g10/pad"]).sum(dim=1)
    assert n_eos.unique().numel() >= 3 and (n_eos < 15).any()      # ... others stop at different steps
    assert gold["eosall/g10/tok"].shape[1] < 15                    # everyone finishes: early exit, ragged lengths
    assert (~gold["eosall/g10/pad"]).sum(dim=1).unique().numel() >= 3
    assert gold["eosall/b3/tok"].shape[2] < 15


def test_noise(gold):
    e0 = synth.synth_embeddings(16, seed=9)
    g = lambda k: gold[f"noise/{k}"]
    assert (orc.noise_gauss_elem(e0, g("gauss_elem/na"), 3.25) - g("gauss_elem/out")).abs().max() < 1e-6
    assert (orc.noise_gauss_vec(e0, g("gauss_vec/na"), g("gauss_vec/ra"), 0.8) - g("gauss_vec/out")).abs().max() < 1e-6
    assert (orc.noise_angle(e0, g("uniform_angle/na"), orc.uniform_angle(g("uniform_angle/ra"), 45.0, 75.0)) - g("uniform_angle/out")).abs().max() < 1e-6
    assert (orc.noise_angle(e0, g("gauss_angle/na"), orc.gauss_angle(g("gauss_angle/ra"), 30.0, 40.0)) - g("gauss_angle/out")).abs().max() < 1e-6
    mix = orc.noise_gauss_elem_uniform_angle(e0, g("mix/na"), g("mix/ra"), g("mix/nb"), g("mix/rb"), 3.25, 45.0, 75.0, 0.5)
    assert (mix - g("mix/out")).abs().max() < 1e-6
    assert ((g("mix/out").norm(dim=1) - 1).abs() < 1e-5).all()


def test_fused_and_spelled_out_maths_agree(monkeypatch):
    """The oracle's default path calls torch's fused CPU functionals (what the reference reaches through
    nn.TransformerEncoder); the spelled-out formulae must give the same numbers."""
    dims = synth.DecoderDims()
    sd = weight_case("eos")
    cfg = orc.cfg_from_state_dict(sd)
    tgt, pad = synth.synth_targets(8, dims, seed=5)
    e = gold_embed()[:8]
    with torch.inference_mode():
        monkeypatch.setattr(orc, "FUSED_OPS", True)
        a, _ = orc.forward_logits(cfg, sd, e, tgt, pad, False)
        monkeypatch.setattr(orc, "FUSED_OPS", False)
        b, _ = orc.forward_logits(cfg, sd, e, tgt, pad, False)
    assert (a - b)[~pad].abs().max() < 1e-4


@pytest.mark.parametrize("tag", ("lively", "eos", "eosall"))
def test_guided_correctness_evaluation(gold, tag):
    """forward(..., guide_targets=...) (embedding_decoder.py:754-760): `correct` compares the target with the arg-max over the ids that
    continue a guide target matching the sequence's own prefix."""
    dims = synth.DecoderDims()
    sd = weight_case(tag)
    cfg = orc.cfg_from_state_dict(sd)
    gt, tgt, pad = guided_eval_case(dims)
    with torch.inference_mode():
        _, _, _, correct = orc.forward_loss(cfg, sd, gold_embed(), tgt, pad, None, guide_targets=gt)
    assert torch.equal(correct, gold[f"{tag}/tfg/correct"])
    assert correct.sum() > gold[f"{tag}/tf/correct"].sum()          # the guide makes the evaluation far more lenient than the plain arg-max


# ----------------------------------------------------------------------------------------------------------------------
# Constructor variants (tests/golden/reference_variants.npz): the oracle restates them too
# ----------------------------------------------------------------------------------------------------------------------
from tests.golden_util import GRAD_CASES, VARIANTS, VARIANTS_PATH, grad_case_inputs, grad_probe, probe_columns, variant_state_dict  # noqa: E402


@pytest.fixture(scope="module")
def vgold():
    return Golden(VARIANTS_PATH)


def _variant_cfg(sd, dims, overrides):
    return orc.cfg_from_state_dict(sd, token_length=dims.token_length, vocab_size=dims.vocab_size,
                                   num_end_loss=overrides.get("num_end_loss", 1), strictly_causal=overrides.get("strictly_causal", False))


@pytest.mark.parametrize("name", list(VARIANTS))
def test_variant_forward_and_greedy(vgold, name):
    dims, overrides = VARIANTS[name]
    sd = variant_state_dict(dims, overrides)
    cfg = _variant_cfg(sd, dims, overrides)
    ls_eps = overrides.get("label_smoothing", 0.0)
    probes = probe_columns(dims.vocab_size)
    embed = synth.synth_embeddings(B_GOLD, dims.embed_dim, seed=1234)
    tgt, pad = synth.synth_targets(B_GOLD, dims, seed=5)
    with torch.inference_mode():
        logits, ls, lb, cor = orc.forward_loss(cfg, sd, embed, tgt, pad, None, label_smoothing=ls_eps)
        o = orc.generate_greedy(cfg, sd, embed, 1.0, 0.0, label_smoothing=ls_eps)
    valid = ~vgold[f"{name}/tf/effpad"]
    assert logits.shape[-1] == dims.vocab_size
    assert (logits[..., probes] - vgold[f"{name}/tf/probes"])[valid].abs().max() < 2e-4
    assert (logits[..., -8:] - vgold[f"{name}/tf/last_cols"])[valid].abs().max() < 2e-4
    assert torch.equal(cor, vgold[f"{name}/tf/correct"])
    assert abs(ls.item() - vgold[f"{name}/tf/loss"][0].item()) < 1e-3 * abs(ls.item())
    assert int(lb) == int(vgold[f"{name}/tf/loss"][1].item())
    assert torch.equal(o["target"], vgold[f"{name}/g10/tok"]) and torch.equal(o["padding"], vgold[f"{name}/g10/pad"])
    assert (o["score"] - vgold[f"{name}/g10/score"]).abs().max() < 2e-4
    assert abs(o["loss_sum"].item() - vgold[f"{name}/g10/loss"][0].item()) < 1e-3 * abs(o["loss_sum"].item())


@pytest.mark.parametrize("name", ["grad_default", "grad_ls01", "grad_multi"])
def test_variant_gradients_of_the_oracle(vgold, name):
    """torch autograd through the oracle reproduces the reference's gradients (which pins the gradient checks of tests/test_gpu_training.py)."""
    dims, overrides, multi = GRAD_CASES[name]
    sd = variant_state_dict(dims, overrides)
    cfg = _variant_cfg(sd, dims, overrides)
    embed, tgt, pad, w = grad_case_inputs(dims, multi)
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k != "causality_mask"}
    A = tgt.shape[0] * max(multi, 1)
    _, loss_sum, _, _ = orc.forward_loss(cfg, leaf, embed, tgt.view(A, -1), pad.view(A, -1), None if w is None else w.view(-1),
                                         label_smoothing=overrides.get("label_smoothing", 0.0))
    loss_sum.backward()
    assert abs(loss_sum.item() - vgold[f"{name}/loss"][0].item()) < 1e-3 * abs(loss_sum.item())
    for i, k in enumerate(sorted(leaf)):
        g = leaf[k].grad.double().flatten()
        idx, proj = grad_probe(i, g.numel())
        r_norm = float(vgold[f"{name}/{k}/norm"])
        assert abs(g.norm().item() - r_norm) <= 1e-3 * r_norm, k
        assert (g[torch.from_numpy(idx)] - vgold[f"{name}/{k}/probe"].double()).abs().max().item() <= 1e-3 * max(vgold[f"{name}/{k}/probe"].abs().max().item(), 1e-9), k
        assert ((torch.from_numpy(proj).double() @ g) - vgold[f"{name}/{k}/proj"].double()).abs().max().item() <= 1e-3 * r_norm, k
