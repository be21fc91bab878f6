"""Guided decoding: the trie built on the host, the CPU oracle against the committed outputs of
the unmodified reference, and (GPU) the CUDA path against both.

Guided decoding has an exact, size-independent invariant: whatever the logits are, every generated sequence must spell one of
the guide targets (up to and including its end token), because only ids that continue a still-matching guide target may ever
be selected (embedding_decoder.py:806-811, :915-917, :942-943)."""
import numpy as np
import pytest
import torch

from novic_b200 import default_decoder, guide, synth
from oracle import novic_oracle as orc
from tests.golden_util import B_GOLD, Golden, gold_embed, weight_case

DEV = "cuda:0"
GUIDE_SETS = {"gflat": (300, 0), "gdeep": (400, 24)}
SCORE_TOL = 0.06 * 15   # |d logit| <= 0.06 per position (tests/test_gpu_parity.py), at most 15 positions per score


def guide_set(name):
    W, pool = GUIDE_SETS[name]
    return synth.synth_guide_targets(W, synth.DecoderDims(), seed=21, first_pool=pool)


def spells_a_guide_target(tok: torch.Tensor, pad: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """tok / pad: N x T generated ids and padding; gt: W x Cmax.  True where the row equals some guide target on every unpadded
    position (the generated row may stop before the target's end token only if it ran out of positions)."""
    T = tok.shape[1]
    eq = (tok.unsqueeze(1) == gt[:, :T].unsqueeze(0)) | pad.unsqueeze(1)        # N x W x T
    return eq.all(dim=2).any(dim=1)


@pytest.fixture(scope="module")
def gold():
    return Golden()


# ------------------------------------------------------------------------------------------------------------
# CPU: trie builder
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(GUIDE_SETS))
def test_trie_children_equal_the_reference_mask_semantics(name):
    """Walking the trie along a prefix must give exactly the ids the reference's W-wide mismatch mask allows."""
    dims = synth.DecoderDims()
    gt = guide_set(name)
    trie = guide.build_trie(gt, dims.token_length - 1, dims.vocab_size)
    off, tok, node = trie.child_off.numpy(), trie.child_tok.numpy(), trie.child_node.numpy()
    assert off[0] == 0 and off[-1] == trie.num_edges and (np.diff(off) >= 0).all()
    rng = np.random.default_rng(3)
    for w in rng.choice(gt.shape[0], size=40, replace=False):
        row = gt[w].numpy()
        n = 0
        mask = np.zeros(gt.shape[0], dtype=bool)
        for c in range(dims.token_length - 1):
            allowed_ref = set(int(t) for t in gt[~torch.from_numpy(mask), c].tolist())
            kids = tok[off[n]:off[n + 1]]
            assert (np.diff(kids) > 0).all()                      # ascending, unique
            assert set(int(t) for t in kids) == allowed_ref
            j = int(np.searchsorted(kids, row[c]))
            assert kids[j] == row[c]
            n = int(node[off[n] + j])
            mask |= gt[:, c].numpy() != row[c]


def test_trie_rejects_bad_input():
    dims = synth.DecoderDims()
    with pytest.raises(ValueError):
        guide.build_trie(torch.full((3, 16), dims.vocab_size, dtype=torch.int64), 15, dims.vocab_size)
    with pytest.raises(ValueError):
        guide.build_trie(torch.zeros((0, 16), dtype=torch.int64), 15, dims.vocab_size)


# ------------------------------------------------------------------------------------------------------------
# CPU: oracle vs committed reference outputs
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag,name", (("lively", "gflat"), ("eos", "gdeep")))
@pytest.mark.parametrize("rname,renorm", (("p", False), ("r", True)))
def test_oracle_guided_vs_reference_outputs(gold, tag, name, rname, renorm):
    sd = weight_case(tag)
    cfg = orc.cfg_from_state_dict(sd)
    gt = guide_set(name)
    with torch.inference_mode():
        o = orc.generate_greedy(cfg, sd, gold_embed(), 0.8, 0.3, guide_targets=gt, guide_renorm=renorm)
        ob = orc.generate_beam(cfg, sd, gold_embed()[:8], 3, 1.0, 0.0, guide_targets=gt, guide_renorm=renorm)
    k = f"{tag}/{name}/greedy_{rname}"
    assert torch.equal(o["target"], gold[f"{k}/tok"]) and torch.equal(o["padding"], gold[f"{k}/pad"])
    assert (o["score"] - gold[f"{k}/score"]).abs().max() < 2e-3
    assert abs(o["loss_sum"].item() - gold[f"{k}/loss"][0].item()) < 2e-3 * abs(o["loss_sum"].item())
    kb = f"{tag}/{name}/beam3_{rname}"
    Tb = ob["target"].shape[2]        # the all-finished early exit is batch-wide: 8 samples may stop before the 32 of the fixture
    assert Tb <= gold[f"{kb}/tok"].shape[2] and not gold[f"{kb}/tok"][:8, :, Tb:].any()
    assert torch.equal(ob["target"], gold[f"{kb}/tok"][:8, :, :Tb]) and torch.equal(ob["padding"], gold[f"{kb}/pad"][:8, :, :Tb])
    assert (ob["score"] - gold[f"{kb}/score"][:8]).abs().max() < 2e-3
    assert spells_a_guide_target(o["target"], o["padding"], gt).all()
    assert spells_a_guide_target(ob["target"].flatten(0, 1), ob["padding"].flatten(0, 1), gt).all()


# ------------------------------------------------------------------------------------------------------------
# GPU: CUDA path
# ------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def models():
    cache = {}

    def get(tag):
        if tag not in cache:
            cache[tag] = default_decoder(synth.DecoderDims(), weight_case(tag)).to(DEV)
        return cache[tag]
    return get


def oracle_guided_scores(cfg, sd, embed, tok, pad, tau, alpha, gt, renorm):
    """Score given sequences the way the guided search does (embedding_decoder.py:915-943): per unpadded position the
    log-softmax of logits / tau, over the whole vocabulary, or - with guide_renorm - over the ids that continue a guide target
    matching the sequence's prefix.  tok / pad: N x T.  Returns N scores (length-normalised when alpha != 0)."""
    N, T = tok.shape
    V = cfg.vocab_size
    full = torch.zeros(N, cfg.token_length, dtype=torch.int64)
    full[:, :T] = tok
    fpad = torch.ones(N, cfg.token_length, dtype=torch.bool)
    fpad[:, :T] = pad
    logits, _ = orc.forward_logits(cfg, sd, embed, full, fpad, only_pred=False)
    logits = logits[:, :T] / tau
    if renorm:
        mismatch = torch.zeros(N, gt.shape[0], dtype=torch.bool)
        for c in range(T):
            gs = orc.guide_score_dense(gt[:, c], mismatch, V, logits.dtype)
            logits[:, c] = logits[:, c] + gs
            mismatch = mismatch | (tok[:, c].unsqueeze(1) != gt[:, c].unsqueeze(0))
    lp = torch.log_softmax(logits, dim=-1).gather(-1, tok.unsqueeze(-1)).squeeze(-1).masked_fill(pad, 0.0)
    s = lp.sum(dim=1)
    if alpha != 0:
        s = s * (~pad).sum(dim=1).clamp(min=1).float().pow(-alpha)
    return s


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ("lively", "eos", "eosall"))
@pytest.mark.parametrize("name", sorted(GUIDE_SETS))
@pytest.mark.parametrize("rname,renorm", (("p", False), ("r", True)))
def test_guided_greedy_vs_reference_outputs(gold, models, tag, name, rname, renorm):
    """A greedy walk that meets a near-tie (closer than the bf16 logit tolerance) may legitimately continue along another
    guide target and end with an unrelated score, so literal equality is required of the large majority of rows, and of
    every row: the exact guide invariant, and a score that is the oracle's score of the very sequence returned."""
    sd = weight_case(tag)
    cfg = orc.cfg_from_state_dict(sd)
    gt = guide_set(name)
    tau, alpha = 0.8, 0.3
    with torch.inference_mode():
        tok, pad, _, ls, lb, score = models(tag).generate(gold_embed().to(DEV), False, True, tau, alpha, None, gt.to(DEV), renorm)
        tok, pad, score = tok.cpu(), pad.cpu(), score.cpu()
        rescored = oracle_guided_scores(cfg, sd, gold_embed(), tok, pad, tau, alpha, gt, renorm)
    k = f"{tag}/{name}/greedy_{rname}"
    rt, rp, rs = gold[f"{k}/tok"], gold[f"{k}/pad"], gold[f"{k}/score"]
    assert spells_a_guide_target(tok, pad, gt).all()                       # exact invariant
    assert (tok[pad] == 0).all()
    assert (rescored - score).abs().max() <= SCORE_TOL / tau
    T = min(tok.shape[1], rt.shape[1])
    same = (tok[:, :T] == rt[:, :T]).all(dim=1) & (pad[:, :T] == rp[:, :T]).all(dim=1)
    assert same.float().mean() >= 0.85, f"only {int(same.sum())}/{B_GOLD} guided greedy rows equal the reference"
    assert (score - rs)[same].abs().max() <= SCORE_TOL / tau
    if same.all() and tok.shape[1] == rt.shape[1]:
        assert int(lb) == int(gold[f"{k}/loss"][1].item())
        assert abs(ls.item() - gold[f"{k}/loss"][0].item()) <= 0.06 * int(lb) + 1e-3 * abs(ls.item())


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ("lively", "eos"))
@pytest.mark.parametrize("name", sorted(GUIDE_SETS))
@pytest.mark.parametrize("H", (3, 10))
@pytest.mark.parametrize("rname,renorm", (("p", False), ("r", True)))
def test_guided_beam_vs_reference_outputs(gold, models, tag, name, H, rname, renorm):
    sd = weight_case(tag)
    cfg = orc.cfg_from_state_dict(sd)
    gt = guide_set(name)
    with torch.inference_mode():
        tok, pad, score = models(tag).generate_beam(gold_embed().to(DEV), H, 1.0, 0.0, None, False, 0.0, gt.to(DEV), renorm)
        tok, pad, score = tok.cpu(), pad.cpu(), score.cpu()
        rescored = oracle_guided_scores(cfg, sd, gold_embed().repeat_interleave(H, dim=0), tok.flatten(0, 1), pad.flatten(0, 1), 1.0, 0.0, gt, renorm).view(-1, H)
    k = f"{tag}/{name}/beam{H}_{rname}"
    rt, rp, rs = gold[f"{k}/tok"], gold[f"{k}/pad"], gold[f"{k}/score"]
    assert spells_a_guide_target(tok.flatten(0, 1), pad.flatten(0, 1), gt).all()   # exact invariant, every beam
    assert (score[:, :-1] >= score[:, 1:]).all()                                   # sorted descending
    assert (tok[pad] == 0).all()
    assert (rescored - score).abs().max() <= SCORE_TOL                             # reported score = oracle's score of that sequence
    T = min(tok.shape[2], rt.shape[2])
    same = (tok[:, :, :T] == rt[:, :, :T]).all(dim=2) & (pad[:, :, :T] == rp[:, :, :T]).all(dim=2)
    assert same[:, 0].float().mean() >= 0.8, f"only {int(same[:, 0].sum())}/{B_GOLD} best beams equal the reference"
    assert same.float().mean() >= 0.6
    assert (score - rs)[same].abs().max() <= SCORE_TOL
    # search quality: the best beam scores like the reference's best beam on (almost) every sample
    assert ((score[:, 0] - rs[:, 0]).abs() <= SCORE_TOL).float().mean() >= 0.9


@pytest.mark.gpu
def test_guided_full_size_properties(models):
    """B = 4096, the reference's default generation config (beam k=10, guided, no renorm; infer.py:55) with a 3000-noun guide:
    every beam spells a guide target, scores are sorted, two runs are bit-identical, guided greedy == guided beam's best
    wherever the beam's top-2 score gap is decisive."""
    dims = synth.DecoderDims()
    model = default_decoder(dims, synth.synth_state_dict(dims, seed=1)).to(DEV)
    embed = synth.synth_embeddings(4096, seed=1234).to(DEV)
    gt = synth.synth_guide_targets(3000, dims, seed=33, first_pool=200)
    gtd = gt.to(DEV)
    with torch.inference_mode():
        tok, pad, score = model.generate_beam(embed, 10, 1.0, 0.0, None, False, 0.0, gtd, False)
        tok2, pad2, score2 = model.generate_beam(embed, 10, 1.0, 0.0, None, False, 0.0, gtd, False)
        g = model.generate(embed, False, True, 1.0, 0.0, None, gtd, False)
    assert torch.equal(tok, tok2) and torch.equal(score, score2)
    tok, pad, score = tok.cpu(), pad.cpu(), score.cpu()
    ok = spells_a_guide_target(tok.flatten(0, 1)[:20000], pad.flatten(0, 1)[:20000], gt)
    assert ok.all()
    assert (score[:, :-1] >= score[:, 1:]).all() and torch.isfinite(score).all()
    assert spells_a_guide_target(g[0].cpu()[:2048], g[1].cpu()[:2048], gt).all()
    # greedy follows the locally best allowed id; a width-10 beam almost always ends at least as well (it can prune the greedy
    # path, so this is a statistical property, not an invariant)
    assert (score[:, 0] >= g[5].cpu() - 1e-3).float().mean() >= 0.95


# ------------------------------------------------------------------------------------------------------------
# generate_all (embedding_decoder.py:986-1079)
# ------------------------------------------------------------------------------------------------------------
ALL_CASES = {   # name: (guide_renorm, vocab prior?, vocab_per_token, vocab_scaler, tau, alpha)
    "plain": (False, False, False, 0.0, 1.0, 0.0), "renorm": (True, False, False, 0.0, 0.8, 0.4),
    "vcount": (True, True, False, 0.7, 1.0, 0.0), "vtoken": (False, True, True, 0.5, 1.0, 0.3),
}


def all_sets():
    dims = synth.DecoderDims()
    gt = synth.synth_guide_targets(200, dims, seed=23, first_pool=24)
    vt = torch.cat((gt[:150], synth.synth_guide_targets(120, dims, seed=24, first_pool=24)))
    return gt, vt


@pytest.mark.parametrize("aname", ("renorm", "vcount"))
def test_oracle_generate_all_vs_reference_outputs(gold, aname):
    renorm, use_vocab, per_token, scaler, tau, alpha = ALL_CASES[aname]
    sd = weight_case("lively")
    cfg = orc.cfg_from_state_dict(sd)
    gt, vt = all_sets()
    with torch.inference_mode():
        o = orc.generate_all(cfg, sd, gold_embed()[:4], gt.shape[0], tau, alpha, gt, renorm, vt if use_vocab else None, per_token, scaler)
    rs = gold[f"lively/all_{aname}/score"][:4]
    fin = torch.isfinite(rs)
    assert (torch.isfinite(o["score"]) == fin).all()
    assert (o["score"] - rs)[fin].abs().max() < 2e-3
    assert torch.equal(o["target"][:, :10], gold[f"lively/all_{aname}/tok10"][:4])
    assert torch.equal(o["padding"][:, :10], gold[f"lively/all_{aname}/pad10"][:4])


def test_generate_all_precompute_matches_oracle_semantics():
    """Host-side pieces of generate_all (no GPU needed): trimmed targets / paddings, vocabulary-prior sums (count-based and
    per-token, +inf when a guide target leaves the vocabulary trie) and the length scale, against a dense restatement."""
    dims = synth.DecoderDims()
    gt, vt = all_sets()
    model = default_decoder(dims, weight_case("lively"))
    for per_token in (False, True):
        targets, pads, trie, vscores, ascale = model.precompute_generate_all(0.4, vt, per_token, 0.7, gt, True)
        W, C = targets.shape
        assert C == 7 and trie is not None and vscores.shape == (1, W) and ascale.shape == (1, W)
        assert torch.equal(pads[:, 1:], (gt[:, :C - 1] == 0).cummax(dim=1).values) and not pads[:, 0].any()
        for w in (0, 17, 151, 199):   # dense check of a few rows
            total = 0.0
            for c in range(C):
                if pads[w, c]:
                    continue
                match = (vt[:, :c] == gt[w, :c]).all(dim=1)
                nxt = vt[match, c]
                if per_token:
                    p = (1.0 / len(set(nxt.tolist()))) if (nxt == gt[w, c]).any() else 0.0
                else:
                    p = float((nxt == gt[w, c]).sum()) / max(int(match.sum()), 1)
                total += float("inf") if p == 0 else -np.log(p) * -1.0
            want = total * 0.7
            got = vscores[0, w].item()
            assert (np.isinf(want) and np.isinf(got)) or abs(want - got) < 1e-4
        n = (C - pads.sum(dim=1)).clamp(min=1).float()
        assert torch.allclose(ascale[0], n.pow(-0.4))


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ("lively", "eos"))
@pytest.mark.parametrize("aname", sorted(ALL_CASES))
def test_generate_all_vs_reference_outputs(gold, models, tag, aname):
    renorm, use_vocab, per_token, scaler, tau, alpha = ALL_CASES[aname]
    gt, vt = all_sets()
    W = gt.shape[0]
    with torch.inference_mode():
        tok, pad, score = models(tag).generate_all(gold_embed()[:16].to(DEV), W, tau, alpha, vt.to(DEV) if use_vocab else None, per_token, scaler,
                                                   gt.to(DEV), renorm)
        tok10, pad10, score10 = models(tag).generate_all(gold_embed()[:16].to(DEV), 10, tau, alpha, vt.to(DEV) if use_vocab else None, per_token,
                                                         scaler, gt.to(DEV), renorm)
    tok, pad, score = tok.cpu(), pad.cpu(), score.cpu()
    rs = gold[f"{tag}/all_{aname}/score"]
    assert tok.shape[:2] == (16, W) and (score[:, :-1] >= score[:, 1:]).all()
    assert torch.equal(tok10.cpu(), tok[:, :10]) and torch.equal(score10.cpu(), score[:, :10]) and torch.equal(pad10.cpu(), pad[:, :10])
    fin = torch.isfinite(rs)
    assert (torch.isfinite(score) == fin).all()                      # guide targets outside the vocabulary score -inf on both sides
    tol = 0.06 / tau * 7                                             # |d logit| <= 0.06 per position, at most 7 positions (C = 7)
    assert (score - rs)[fin].abs().max() <= tol                      # the whole ranking, rank by rank
    # the top-10 sets agree up to near-ties: every returned target is a guide target, and most are the reference's
    assert spells_a_guide_target(tok[:, :10].flatten(0, 1), pad[:, :10].flatten(0, 1), gt).all()
    r10 = gold[f"{tag}/all_{aname}/tok10"]
    agree = sum(len({tuple(r) for r in tok[b, :10].tolist()} & {tuple(r) for r in r10[b].tolist()}) for b in range(16))
    assert agree >= 0.9 * 160
    assert (tok[:, 0] == r10[:, 0]).all(dim=1).float().mean() >= 0.8
