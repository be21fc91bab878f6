"""Beam search with a vocabulary prior (embedding_decoder.py:881-891, :924-936, :972-975).

The reference divides the vocabulary prior out of every candidate's score: with `vocab_targets` (Z x Cmax nouns) and
`vocab_scaler` s, a continuation `t` of prefix `p` costs an extra  -s * log p_vocab(t | p), where p_vocab is the fraction of the
vocabulary nouns matching `p` that continue with `t` (or, per-token, uniform over the distinct continuations); ids no
vocabulary noun continues with score -inf.  Three set-ups exist: vocabulary alone (it then also restricts decoding to the
vocabulary nouns), vocabulary == guide targets, and a vocabulary that differs from the guide.

CPU: the oracle against committed outputs of the unmodified reference; the host-side per-edge prior against a dense
restatement.  GPU: the CUDA path against the same fixtures, with the exact invariants (every beam spells a noun of the
restricting set; reported score = the oracle's score of the returned sequence)."""
import numpy as np
import pytest
import torch

from novic_b200 import default_decoder, guide, synth
from oracle import novic_oracle as orc
from tests.golden_util import B_GOLD, Golden, gold_embed, weight_case
from tests.test_guided import SCORE_TOL, spells_a_guide_target

DEV = "cuda:0"

# name: (H, guided?, vocabulary = "v" (differs from guide) | "g" (equal to the guide), per_token, scaler, guide_renorm, tau, alpha)
CASES = {
    "vonly_c": (3, False, "v", False, 0.6, False, 1.0, 0.0), "vonly_t": (10, False, "v", True, 0.4, False, 0.9, 0.3),
    "vguide_c": (10, True, "g", False, 0.5, False, 1.0, 0.0), "vguide_t": (3, True, "g", True, 0.8, True, 1.0, 0.0),
    "vdiff_c": (10, True, "v", False, 0.5, True, 1.0, 0.0), "vdiff_t": (3, True, "v", True, 0.3, False, 1.2, 0.5),
}


def sets():
    dims = synth.DecoderDims()
    gt = synth.synth_guide_targets(400, dims, seed=21, first_pool=24)
    vt = torch.cat((gt[:250], synth.synth_guide_targets(200, dims, seed=25, first_pool=24)))
    return gt, vt


def case_args(name):
    H, guided, vsel, per_token, scaler, renorm, tau, alpha = CASES[name]
    gt, vt = sets()
    g = gt if guided else None
    v = gt.clone() if vsel == "g" else vt
    return H, g, v, per_token, scaler, renorm, tau, alpha


@pytest.fixture(scope="module")
def gold():
    return Golden()


def dense_prior_logp(prefix, vocab, per_token):
    """log p_vocab(. | prefix) over all ids, by brute force (embedding_decoder.py:925-934); prefix: list of ids."""
    c = len(prefix)
    match = (vocab[:, :c] == torch.tensor(prefix, dtype=torch.int64).view(1, -1)).all(dim=1) if c else torch.ones(vocab.shape[0], dtype=torch.bool)
    nxt = vocab[match, c]
    out = {}
    ids, counts = np.unique(nxt.numpy(), return_counts=True)
    for i, n in zip(ids.tolist(), counts.tolist()):
        out[i] = -np.log(len(ids)) if per_token else np.log(n / int(match.sum()))
    return out


# ------------------------------------------------------------------------------------------------------------
# CPU
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("per_token", (False, True))
@pytest.mark.parametrize("mode", ("vonly", "vguide", "vdiff"))
def test_prior_bias_per_edge_matches_dense_restatement(mode, per_token):
    dims = synth.DecoderDims()
    G, V = dims.token_length - 1, dims.vocab_size
    gt, vt = sets()
    scaler = 0.7
    if mode == "vonly":
        trie, bias = guide.prior_bias(None, vt, False, per_token, scaler, G, V)
        walk_set, vocab = vt, vt
    elif mode == "vguide":
        trie, bias = guide.prior_bias(guide.build_trie(gt, G, V), gt, True, per_token, scaler, G, V)
        walk_set, vocab = gt, gt
    else:
        trie, bias = guide.prior_bias(guide.build_trie(gt, G, V), vt, False, per_token, scaler, G, V)
        walk_set, vocab = gt, vt
    assert bias.dtype == torch.float32 and bias.shape == (trie.num_edges,)
    off, tok, node = trie.child_off.numpy(), trie.child_tok.numpy(), trie.child_node.numpy()
    rng = np.random.default_rng(3)
    seen_inf = False
    for w in rng.choice(walk_set.shape[0], size=25, replace=False):      # walk 25 nouns through the trie, checking every edge on the way
        n, prefix = 0, []
        for c in range(G):
            t = int(walk_set[w, c])
            want_all = dense_prior_logp(prefix, vocab, per_token)
            e0, e1 = off[n], off[n + 1]
            for e in range(e0, e1):                                      # every child edge of the node, not only the one taken
                got = bias[e].item()
                if int(tok[e]) in want_all:
                    assert abs(got - (-scaler * want_all[int(tok[e])])) < 1e-5
                else:
                    assert got == float("-inf")
                    seen_inf = True
            e = e0 + int(np.searchsorted(tok[e0:e1], t))
            assert e < e1 and tok[e] == t
            n = int(node[e])
            prefix.append(t)
    assert seen_inf == (mode == "vdiff")          # only a guide that leaves the vocabulary has impossible edges


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_vocab_prior_beam_vs_reference_outputs(gold, name):
    H, g, v, per_token, scaler, renorm, tau, alpha = case_args(name)
    tag = "eos" if name.endswith("_c") else "lively"
    sd = weight_case(tag)
    cfg = orc.cfg_from_state_dict(sd)
    n = 6
    with torch.inference_mode():
        o = orc.generate_beam(cfg, sd, gold_embed()[:n], H, tau, alpha, guide_targets=g, guide_renorm=renorm, vocab_targets=v,
                              vocab_per_token=per_token, vocab_scaler=scaler)
    rt, rp, rs = gold[f"{tag}/{name}/tok"][:n], gold[f"{tag}/{name}/pad"][:n], gold[f"{tag}/{name}/score"][:n]
    T = o["target"].shape[2]             # the all-finished early exit is batch-wide: 6 samples may stop before the fixture's 32
    assert T <= rt.shape[2] and not rt[:, :, T:].any()
    live = torch.isfinite(rs)            # beams beyond the number of reachable nouns are -inf fillers with arbitrary ids
    assert (torch.isfinite(o["score"]) == live).all()
    assert torch.equal(o["target"][live], rt[:, :, :T][live]) and torch.equal(o["padding"][live], rp[:, :, :T][live])
    assert (o["score"] - rs)[live].abs().max() < 2e-3


# ------------------------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def models():
    cache = {}

    def get(tag):
        if tag not in cache:
            cache[tag] = default_decoder(synth.DecoderDims(), weight_case(tag)).to(DEV)
        return cache[tag]
    return get


def oracle_prior_scores(cfg, sd, embed, tok, pad, tau, alpha, g, renorm, v, per_token, scaler):
    """Score of given sequences under the reference's beam objective (embedding_decoder.py:915-943): per unpadded position
    log_softmax(logits / tau [restricted to the guide's continuations with guide_renorm]) - scaler * log p_vocab(token | prefix)."""
    N, T = tok.shape
    V = cfg.vocab_size
    full = torch.zeros(N, cfg.token_length, dtype=torch.int64)
    full[:, :T] = tok
    fpad = torch.ones(N, cfg.token_length, dtype=torch.bool)
    fpad[:, :T] = pad
    logits, _ = orc.forward_logits(cfg, sd, embed, full, fpad, only_pred=False)
    logits = logits[:, :T] / tau
    if renorm and g is not None:
        mismatch = torch.zeros(N, g.shape[0], dtype=torch.bool)
        for c in range(T):
            logits[:, c] = logits[:, c] + orc.guide_score_dense(g[:, c], mismatch, V, logits.dtype)
            mismatch = mismatch | (tok[:, c].unsqueeze(1) != g[:, c].unsqueeze(0))
    lp = torch.log_softmax(logits, dim=-1).gather(-1, tok.unsqueeze(-1)).squeeze(-1)
    prior = torch.zeros(N, T)
    cache = {}
    for i in range(N):
        for c in range(T):
            if pad[i, c]:
                continue
            key = tuple(tok[i, :c].tolist())
            if key not in cache:
                cache[key] = dense_prior_logp(list(key), v, per_token)
            prior[i, c] = cache[key].get(int(tok[i, c]), float("-inf"))
    s = (lp - scaler * prior).masked_fill(pad, 0.0).sum(dim=1)
    if alpha != 0:
        s = s * (~pad).sum(dim=1).clamp(min=1).float().pow(-alpha)
    return s


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ("lively", "eos"))
@pytest.mark.parametrize("name", sorted(CASES))
def test_vocab_prior_beam_vs_reference_outputs(gold, models, tag, name):
    H, g, v, per_token, scaler, renorm, tau, alpha = case_args(name)
    sd = weight_case(tag)
    cfg = orc.cfg_from_state_dict(sd)
    with torch.inference_mode():
        tok, pad, score = models(tag).generate_beam(gold_embed().to(DEV), H, tau, alpha, v.to(DEV), per_token, scaler,
                                                    None if g is None else g.to(DEV), renorm)
        tok, pad, score = tok.cpu(), pad.cpu(), score.cpu()
    rt, rp, rs = gold[f"{tag}/{name}/tok"], gold[f"{tag}/{name}/pad"], gold[f"{tag}/{name}/score"]
    live = torch.isfinite(score)
    assert (live == torch.isfinite(rs)).all()                                      # same number of reachable nouns per sample
    assert (score[:, :-1] >= score[:, 1:])[live[:, 1:]].all()                      # sorted descending
    assert (tok[pad] == 0).all()
    restrict = g if g is not None else v                                           # exact invariant: every live beam spells a noun
    assert spells_a_guide_target(tok[live], pad[live], restrict).all()
    if g is not None:                                                              # ... that is also in the vocabulary (else its score is -inf)
        assert spells_a_guide_target(tok[live], pad[live], v).all()
    with torch.inference_mode():
        rescored = oracle_prior_scores(cfg, sd, gold_embed().repeat_interleave(H, dim=0)[live.flatten()], tok[live], pad[live], tau, alpha, g,
                                       renorm, v, per_token, scaler)
    assert (rescored - score[live]).abs().max() <= SCORE_TOL / tau                 # reported score = oracle's score of that sequence
    T = min(tok.shape[2], rt.shape[2])
    same = (tok[:, :, :T] == rt[:, :, :T]).all(dim=2) & (pad[:, :, :T] == rp[:, :, :T]).all(dim=2)
    assert same[:, 0].float().mean() >= 0.8, f"only {int(same[:, 0].sum())}/{B_GOLD} best beams equal the reference"
    assert same[live].float().mean() >= 0.6
    assert (score - rs)[same & live].abs().max() <= SCORE_TOL / tau
    assert ((score[:, 0] - rs[:, 0]).abs() <= SCORE_TOL / tau).float().mean() >= 0.9


@pytest.mark.gpu
def test_vocab_prior_arguments(models):
    m = models("lively")
    gt, vt = sets()
    e = gold_embed()[:4].to(DEV)
    with torch.inference_mode():
        a = m.generate_beam(e, 3, 1.0, 0.0, vt.to(DEV), False, 0.0, None, False)        # scaler 0 = no prior (embedding_decoder.py:881)
        b = m.generate_beam(e, 3, 1.0, 0.0, None, False, 0.0, None, False)
        assert torch.equal(a[0], b[0]) and torch.equal(a[2], b[2])
        with pytest.raises(ValueError):
            m.generate_beam(e, 3, 1.0, 0.0, vt.to(DEV), False, -0.5, None, False)
