"""CPU oracle for the NOVIC object-noun decoder hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch CPU restatement (torch CPU tensors, explicit maths) of the algorithm the
reference implements in `embedding_decoder.py` / `embedding_noise.py`.  It exists to *check* the CUDA
product path; only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs
of `bench.py` may import it.  Nothing under `novic_b200/` imports it and the product path has no CPU
fallback.

Parity status: PINNED.  `oracle/validate_vs_reference.py` runs this restatement against the reference's own
classes imported from /root/reference (possible only in the build container) and
`oracle/make_golden.py` stores outputs of the *reference itself* under `tests/golden/`; the CPU test suite
checks this oracle against those committed vectors (tests/test_oracle_golden.py).

Every function cites the reference lines it restates (paths relative to /root/reference).
The schedule deliberately mirrors the reference (no KV cache: every decode step re-runs the whole
prefix+tokens sequence, `embedding_decoder.py:792-798`), so that timing this oracle on host cores is a
fair "port" of the reference's CPU inference path.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Optional

import numpy as np
import torch

NEG_INF = float("-inf")


@dataclasses.dataclass(frozen=True)
class OracleCfg:
    """Architecture constants (config/train.yaml:224-308 defaults)."""
    embed_dim: int = 1024      # F
    hidden_dim: int = 512      # E
    ffn_dim: int = 128         # K = E * feedfwd_scale (1/4)
    num_layers: int = 6        # L
    num_heads: int = 8
    prefix_len: int = 4        # P = mlp_seq_len
    vocab_size: int = 6912     # V
    token_length: int = 16     # Cmax (includes the trailing end token)
    ln_eps: float = 1e-5
    num_end_loss: int = 1
    strictly_causal: bool = False   # embedding_decoder.py:652: without it the P x P prefix block is bidirectional

    @property
    def max_seq_len(self) -> int:  # embedding_decoder.py:648
        return self.prefix_len + self.token_length - 1

    @property
    def gen_len(self) -> int:  # embedding_decoder.py:782
        return self.token_length - 1


def cfg_from_state_dict(sd: dict, token_length: Optional[int] = None, num_heads: int = 8, vocab_size: Optional[int] = None,
                        num_end_loss: int = 1, strictly_causal: bool = False) -> OracleCfg:
    """vocab_size: pass the real V when the tied matrix carries vocab_quant rows (embedding_decoder.py:642-645, :726-727)."""
    E = sd["logits_linear.weight"].shape[1]
    V = sd["logits_linear.weight"].shape[0] if vocab_size is None else vocab_size
    F = sd["embed_mlp.mlp.0.weight"].shape[1]
    P = sd["embed_mlp.mlp.0.weight"].shape[0] // E
    L = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.layers."))
    K = sd["transformer.layers.0.linear1.weight"].shape[0]
    S = sd["pos_embedding.embedding.weight"].shape[0]
    Cmax = S - P + 1 if token_length is None else token_length
    return OracleCfg(embed_dim=F, hidden_dim=E, ffn_dim=K, num_layers=L, num_heads=num_heads, prefix_len=P,
                     vocab_size=V, token_length=Cmax, num_end_loss=num_end_loss, strictly_causal=strictly_causal)


# ----------------------------------------------------------------------------------------------------
# Building blocks
# ----------------------------------------------------------------------------------------------------

# The reference reaches torch's fused CPU kernels (native_layer_norm, gelu, scaled_dot_product_attention) through
# nn.TransformerEncoder.  FUSED_OPS=True makes the oracle call the same torch functionals, so that timing it is a
# fair port of the reference's CPU path; FUSED_OPS=False spells the maths out (tests check the two agree).
FUSED_OPS = True


def _layer_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    # nn.LayerNorm(bias=False): biased variance, eps inside the sqrt (built at embedding_decoder.py:309-327)
    if FUSED_OPS:
        return torch.nn.functional.layer_norm(x, (x.shape[-1],), w, None, eps)
    mu = x.mean(dim=-1, keepdim=True)
    var = (x - mu).square().mean(dim=-1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + eps) * w


def _gelu_erf(x: torch.Tensor) -> torch.Tensor:
    # exact GELU (utils.py:107 -> torch.nn.functional.gelu default)
    if FUSED_OPS:
        return torch.nn.functional.gelu(x)
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def attention_bias(cfg: OracleCfg, S: int, dtype: torch.dtype) -> torch.Tensor:
    """S x S additive mask: causal, except that the P x P prefix block is fully visible
    (embedding_decoder.py:651-654)."""
    q = torch.arange(S).unsqueeze(1)
    k = torch.arange(S).unsqueeze(0)
    visible = (k <= q) if cfg.strictly_causal else ((k <= q) | ((q < cfg.prefix_len) & (k < cfg.prefix_len)))
    bias = torch.zeros(S, S, dtype=dtype)
    bias.masked_fill_(~visible, NEG_INF)
    return bias


def key_padding_bias(cfg: OracleCfg, padding: torch.Tensor, S: int, dtype: torch.dtype) -> tuple[torch.Tensor, torch.Tensor]:
    """Restates embedding_decoder.py:696-712.  padding is A x C bool.  Returns (A x S additive key bias,
    A x C effective target padding)."""
    A, C = padding.shape
    expand = cfg.prefix_len + cfg.num_end_loss - 2
    keep = C - cfg.num_end_loss + 1
    if expand < 1:
        seq_pad = padding
        eff = padding
    else:
        if keep <= 1:
            seq_pad = padding[:, 0:1].expand(-1, S)
        else:
            seq_pad = torch.cat((padding[:, 0:1].expand(-1, expand), padding[:, :keep]), dim=1)
        eff = seq_pad[:, -C:]
    bias = torch.zeros(A, S, dtype=dtype)
    if S > 1:
        bias[:, 1:].masked_fill_(seq_pad[:, 1:], NEG_INF)  # sequence position 0 is never masked (:711-712)
    return bias, eff


class DropMasks:
    """The dropout masks of novic_b200's training step, replayed on the CPU (test infrastructure): the product draws them from a
    counter-based hash of (seed, site, element index) - csrc/ptx.cuh:drop_hash - so that its backward pass can regenerate them; the
    reference uses torch's dropout stream, which no other implementation can reproduce, so parity under dropout is defined as
    "same function given the same masks".  site = layer * 8 + kind; kinds as in ptx.cuh."""
    INPUT, ATTN, BRANCH1, FFN, BRANCH2 = 0, 1, 2, 3, 4

    def __init__(self, p_input: float, p_layer: float, seed: int):
        self.seed32 = np.uint32(((seed ^ (seed >> 32)) & 0xFFFFFFFF))
        self.p = {"input": float(p_input), "layer": float(p_layer)}

    def factor(self, which: str, layer: int, kind: int, shape: tuple) -> Optional[torch.Tensor]:
        """1 / (1 - p) where the element is kept, 0 where it is dropped; element index = C-order position in `shape`."""
        p = self.p[which]
        thresh = np.uint32(int(p * 16777216.0 + 0.5))
        if thresh == 0:
            return None
        n = int(np.prod(shape))
        with np.errstate(over="ignore"):
            x = np.arange(n, dtype=np.uint32) * np.uint32(0x9E3779B1) + self.seed32 + np.uint32(layer * 8 + kind) * np.uint32(0x85EBCA77)
            x ^= x >> np.uint32(16); x *= np.uint32(0x7FEB352D); x ^= x >> np.uint32(15); x *= np.uint32(0x846CA68B); x ^= x >> np.uint32(16)
        keep = (x >> np.uint32(8)) >= thresh
        scale = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
        return torch.from_numpy(np.where(keep, scale, np.float32(0.0)).astype(np.float32).reshape(shape))


def transformer_stack(cfg: OracleCfg, sd: dict, x: torch.Tensor, bias: torch.Tensor, drop: Optional[DropMasks] = None) -> torch.Tensor:
    """Pre-LN encoder stack + final LayerNorm (nn.TransformerEncoder built at embedding_decoder.py:309-327,
    called at :714).  x: A x S x E, bias: broadcastable to A x 1 x S x S (additive).  drop: training-mode dropout with
    explicit masks (attention probabilities, both residual branches, the activated feed-forward rows)."""
    A, S, E = x.shape
    H = cfg.num_heads
    d = E // H
    scale = 1.0 / math.sqrt(d)

    def dropped(t, layer, kind):
        f = None if drop is None else drop.factor("layer", layer, kind, tuple(t.shape))
        return t if f is None else t * f

    for l in range(cfg.num_layers):
        p = f"transformer.layers.{l}."
        h = _layer_norm(x, sd[p + "norm1.weight"], cfg.ln_eps)
        qkv = h @ sd[p + "self_attn.in_proj_weight"].t()
        q, k, v = qkv.split(E, dim=-1)
        q = q.view(A, S, H, d).transpose(1, 2)
        k = k.view(A, S, H, d).transpose(1, 2)
        v = v.view(A, S, H, d).transpose(1, 2)
        if FUSED_OPS and drop is None:
            o = torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=bias.expand(A, 1, S, S) if bias.shape[0] != A else bias)
        else:
            att = torch.softmax((q @ k.transpose(-1, -2)) * scale + bias, dim=-1)
            o = dropped(att, l, DropMasks.ATTN) @ v
        o = o.transpose(1, 2).reshape(A, S, E)
        x = x + dropped(o @ sd[p + "self_attn.out_proj.weight"].t(), l, DropMasks.BRANCH1)
        h = _layer_norm(x, sd[p + "norm2.weight"], cfg.ln_eps)
        h = dropped(_gelu_erf(h @ sd[p + "linear1.weight"].t()), l, DropMasks.FFN)
        x = x + dropped(h @ sd[p + "linear2.weight"].t(), l, DropMasks.BRANCH2)
    return _layer_norm(x, sd["transformer.norm.weight"], cfg.ln_eps)


def forward_logits(cfg: OracleCfg, sd: dict, embed: torch.Tensor, target: Optional[torch.Tensor],
                   padding: Optional[torch.Tensor], only_pred: bool, drop: Optional[DropMasks] = None) -> tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Restates PrefixedIterDecoder.forward up to the logits (embedding_decoder.py:659-727).
    embed B x F; target A x C (A = B*M, sequences of one embedding adjacent, i.e. B-before-M) or None;
    padding A x C bool or None.  Returns (A x T x V logits, A x T effective padding or None)."""
    dtype = embed.dtype
    B = embed.shape[0]
    P, E = cfg.prefix_len, cfg.hidden_dim
    Wt = sd["logits_linear.weight"]
    x = torch.nn.functional.normalize(embed, dim=-1) @ sd["embed_mlp.mlp.0.weight"].t()  # :1276
    x = x.view(B, P, E)
    if target is not None:
        A = target.shape[0]
        if A != B:
            x = x.repeat_interleave(A // B, dim=0)  # :674
        if target.shape[1] > 1:
            x = torch.cat((x, Wt[target[:, :-1]]), dim=1)  # tied token embedding, :692, utils.py:65-68
    S = x.shape[1]
    x = x + sd["pos_embedding.embedding.weight"][:S]  # :1297 (dropout is identity in eval)
    if drop is not None:
        f = drop.factor("input", 0, DropMasks.INPUT, tuple(x.shape))
        x = x if f is None else x * f
    bias = attention_bias(cfg, S, dtype).view(1, 1, S, S)
    eff_pad = None
    if padding is not None:
        kb, eff_pad = key_padding_bias(cfg, padding, S, dtype)
        bias = bias + kb.view(-1, 1, 1, S)
    x = transformer_stack(cfg, sd, x, bias, drop)
    if only_pred:
        x = x[:, -1:, :]
        if eff_pad is not None:
            eff_pad = eff_pad[:, -1:]
    else:
        x = x[:, P - 1:, :]
    return x @ Wt[:cfg.vocab_size].t(), eff_pad    # vocab_quant rows (if any) are sliced away, :726-727


def forward_loss(cfg: OracleCfg, sd: dict, embed: torch.Tensor, target: torch.Tensor, padding: Optional[torch.Tensor],
                 weight: Optional[torch.Tensor], label_smoothing: float = 0.0, drop: Optional[DropMasks] = None,
                 guide_targets: Optional[torch.Tensor] = None):
    """Teacher-forced forward with loss/correct (embedding_decoder.py:729-761), only_pred=False.
    Returns (logits A x C x V, loss_sum, loss_basis, correct A x C)."""
    if weight is not None:  # :681-685
        wpad = (weight == 0).unsqueeze(1)
        padding = wpad.expand_as(target) if padding is None else (padding | wpad)
    logits, eff_pad = forward_logits(cfg, sd, embed, target, padding, only_pred=False, drop=drop)
    A, C, V = logits.shape
    tgt = target if eff_pad is None else target.masked_fill(eff_pad, -1)
    logp = torch.log_softmax(logits, dim=-1)
    valid = tgt >= 0
    picked = logp.gather(2, tgt.clamp(min=0).unsqueeze(2)).squeeze(2)
    nll = -picked
    if label_smoothing != 0.0:
        nll = (1.0 - label_smoothing) * nll + label_smoothing * (-logp.mean(dim=-1))
    nll = nll.masked_fill(~valid, 0.0)
    if weight is None:
        loss_sum = nll.sum()
        loss_basis = valid.sum() if eff_pad is not None else torch.tensor(tgt.numel())
    else:
        loss_sum = (weight * nll.sum(dim=1)).sum()
        n = valid.sum(dim=1) if eff_pad is not None else torch.full((A,), C)
        loss_basis = (weight * n.to(weight.dtype)).sum()
    if guide_targets is None:
        pred = logits.argmax(dim=2)
    else:                                                                       # embedding_decoder.py:754-760
        W = guide_targets.shape[0]
        gT = guide_targets.t()                                                  # Cmax x W
        mism = torch.cat((torch.zeros(A, 1, W, dtype=torch.bool),
                          (target[:, :C - 1, None] != gT[None, :C - 1, :]).cummax(dim=1).values), dim=1)   # A x C x W
        gs = torch.full((A, C, V + 1), NEG_INF, dtype=logits.dtype).scatter_(2, gT[None, :C, :].expand(A, -1, -1).masked_fill(mism, V), 0.0)[:, :, :-1]
        pred = (gs + logits).argmax(dim=2)
    correct = (pred == tgt)
    return logits, loss_sum, loss_basis, correct


# ----------------------------------------------------------------------------------------------------
# Greedy decode (embedding_decoder.py:779-850, unguided)
# ----------------------------------------------------------------------------------------------------

def guide_score_dense(guide_tok: torch.Tensor, guide_mask: torch.Tensor, V: int, dtype: torch.dtype) -> torch.Tensor:
    """0 where a still-matching guide target continues with that token id, -inf elsewhere
    (embedding_decoder.py:807, :916-917).  guide_tok: W ids of the current position, guide_mask: ... x W mismatch flags."""
    idx = guide_tok.expand(guide_mask.shape).masked_fill(guide_mask, V)
    full = torch.full(guide_mask.shape[:-1] + (V + 1,), NEG_INF, dtype=dtype)
    return full.scatter_(-1, idx, 0.0)[..., :-1]


def generate_greedy(cfg: OracleCfg, sd: dict, embed: torch.Tensor, temperature: float = 1.0, length_alpha: float = 0.0,
                    sample_weight: Optional[torch.Tensor] = None, early_exit: bool = True,
                    guide_targets: Optional[torch.Tensor] = None, guide_renorm: bool = False, label_smoothing: float = 0.0):
    """Returns dict(target B x T int64, padding B x T bool, logits B x T x V, loss_sum, loss_basis, score B).
    guide_targets (W x Cmax, :802-813): the generated ids must spell one of the guide targets."""
    B = embed.shape[0]
    G = cfg.gen_len
    tok = torch.zeros(B, G, dtype=torch.int64)
    pad = torch.zeros(B, G, dtype=torch.bool)
    done = torch.zeros(B, dtype=torch.bool)
    step_logits = []
    guide_scores = []
    guide_mask = torch.zeros(B, guide_targets.shape[0], dtype=torch.bool) if guide_targets is not None else None  # :788
    T = G
    for c in range(1, G + 1):
        if c > 1:
            pad[:, c - 1] = done  # :797 - the EOS itself is not padding, the positions after it are
        logits, _ = forward_logits(cfg, sd, embed, tok[:, :c], done.unsqueeze(1).expand(-1, c), only_pred=True)
        logits = logits[:, 0, :]
        step_logits.append(logits)
        if guide_targets is not None:
            gt = guide_targets[:, c - 1]
            gs = guide_score_dense(gt, guide_mask, cfg.vocab_size, logits.dtype)  # :806-807
            guide_scores.append(gs)
            nxt = (gs + logits).argmax(dim=1)                                      # :810 (no end-token ban in the guided branch)
            guide_mask = guide_mask | (nxt.unsqueeze(1) != gt.unsqueeze(0))         # :811
        elif c == 1:
            nxt = logits[:, 1:].argmax(dim=1) + 1  # first token may not be EOS (:804)
        else:
            nxt = logits.argmax(dim=1)
        tok[:, c - 1] = nxt
        done = done | (nxt == 0)
        if early_exit and bool(done.all()):  # :817-820
            T = c
            break
    tok = tok[:, :T].clone()
    pad = pad[:, :T].clone()
    seq_logits = torch.stack(step_logits, dim=1)
    tok.masked_fill_(pad, 0)  # :824
    score_logits = seq_logits / temperature
    if guide_targets is not None and guide_renorm:
        score_logits = score_logits + torch.stack(guide_scores, dim=1)  # :829-830: renormalise over the allowed ids
    logp_t = torch.log_softmax(score_logits, dim=2)
    score = logp_t.gather(2, tok.unsqueeze(2)).squeeze(2).masked_fill(pad, 0.0).sum(dim=1)  # :831-834
    length = (T - pad.sum(dim=1)).to(score.dtype)
    if length_alpha != 0:
        score = score * length.clamp(min=1).pow(-length_alpha)  # :836
    logp = torch.log_softmax(seq_logits, dim=2)
    nll = -logp.gather(2, tok.unsqueeze(2)).squeeze(2)
    if label_smoothing != 0.0:                                           # F.cross_entropy(label_smoothing=...), :840 / :844
        nll = (1.0 - label_smoothing) * nll + label_smoothing * (-logp.mean(dim=-1))
    nll = nll.masked_fill(pad, 0.0)
    if sample_weight is None:
        loss_sum = nll.sum()
        loss_basis = (~pad).sum()
    else:
        loss_sum = (sample_weight * nll.sum(dim=1)).sum()
        loss_basis = (sample_weight * length).sum()
    return dict(target=tok, padding=pad, logits=seq_logits, loss_sum=loss_sum, loss_basis=loss_basis, score=score)


# ----------------------------------------------------------------------------------------------------
# Beam search (embedding_decoder.py:852-984, unguided, no vocab prior)
# ----------------------------------------------------------------------------------------------------

def generate_beam(cfg: OracleCfg, sd: dict, embed: torch.Tensor, topk: int, temperature: float = 1.0,
                  length_alpha: float = 0.0, early_exit: bool = True, guide_targets: Optional[torch.Tensor] = None,
                  guide_renorm: bool = False, vocab_targets: Optional[torch.Tensor] = None, vocab_per_token: bool = False,
                  vocab_scaler: float = 0.0):
    """Returns dict(target B x H x T, padding B x H x T, score B x H sorted descending, margin B).

    `margin` is test metadata, not part of the reference's outputs: per sample, the smallest gap seen at any step
    between adjacent entries of the top-(H+1) ranking values, i.e. how far the search was from pruning or ordering
    its candidates differently.  A lower-precision implementation can only be required to reproduce the beams of
    samples whose margin exceeds its accumulated score tolerance."""
    B, H, G, V = embed.shape[0], topk, cfg.gen_len, cfg.vocab_size
    dtype = embed.dtype
    tok = torch.zeros(B, H, G, dtype=torch.int64)
    pad = torch.ones(B, H, G, dtype=torch.bool)
    pad[:, 0, 0] = False                                  # one live empty candidate per sample (:862)
    score = torch.full((B, H), NEG_INF, dtype=dtype)
    score[:, 0] = 0.0                                     # :864
    score_normed = score.clone()
    seq_len = torch.zeros(B, H, dtype=dtype)
    seq_len[:, 0] = 1.0                                   # :899
    T = G
    margin = torch.full((B,), float("inf"), dtype=dtype)
    guide_mask = None
    if guide_targets is not None:                                             # :873-878
        guide_mask = torch.ones(B, H, guide_targets.shape[0], dtype=torch.bool)
        guide_mask[:, 0, :] = False
    vocab_on = vocab_targets is not None and vocab_scaler != 0                # :880-891
    vocab_is_guide = vocab_on and guide_targets is not None and (vocab_targets is guide_targets or torch.equal(vocab_targets, guide_targets))
    vocab_mask = None
    if vocab_on and not vocab_is_guide:
        vocab_mask = torch.ones(B, H, vocab_targets.shape[0], dtype=torch.bool)
        vocab_mask[:, 0, :] = False
    for c in range(1, G + 1):
        cur_tok = tok[:, :, :c].reshape(B * H, c)
        cur_pad = pad[:, :, :c].reshape(B * H, c)
        logits, lpad = forward_logits(cfg, sd, embed, cur_tok, cur_pad, only_pred=True)
        logits = logits.view(B, H, V) / temperature
        finished = lpad.view(B, H, 1)
        logits[:, :, 1:] = logits[:, :, 1:].masked_fill(finished, NEG_INF)  # finished candidates extend with EOS at zero cost (:913)
        gs = None
        if guide_targets is not None:
            gs = guide_score_dense(guide_targets[:, c - 1], guide_mask, V, logits.dtype)  # :916-917
            gs[:, :, :1] = gs[:, :, :1].masked_fill(finished, 0.0)                        # :918
            if guide_renorm:
                logits = logits + gs                                                        # :920
        cand = torch.log_softmax(logits, dim=2)                             # :922
        if vocab_on:                                                        # :924-936: divide out the vocabulary prior
            if vocab_is_guide:
                vidx = guide_targets[:, c - 1].expand(B, H, -1).masked_fill(guide_mask, V)
            else:
                vidx = vocab_targets[:, c - 1].expand(B, H, -1).masked_fill(vocab_mask, V)
            Z = vidx.shape[2]
            if vocab_per_token:
                vp = torch.zeros(B, H, V + 1, dtype=dtype).scatter_(2, vidx, 1.0)[:, :, :-1]
                vp = vp / vp.sum(dim=2, keepdim=True)
            else:
                cnt = torch.zeros(B, H, V + 1, dtype=dtype).scatter_add_(2, vidx, torch.ones(B, H, Z, dtype=dtype))
                vp = cnt[:, :, :-1] / (Z - cnt[:, :, -1:])
            vlp = vp.log().nan_to_num(nan=float("inf"), neginf=float("inf"), posinf=float("inf"))
            vlp[:, :, :1] = vlp[:, :, :1].masked_fill(finished, 0.0)
            cand = cand - vocab_scaler * vlp
        cand = cand + score.unsqueeze(2)                                    # :938
        if c == 1:
            cand[:, 0, 0] = NEG_INF                                         # :940
        if gs is not None and not guide_renorm:
            cand = cand + gs                                                # :943
        flat = cand.view(B, H * V)
        if length_alpha == 0:
            ranked = flat
            best, idx = torch.topk(flat, k=H, dim=1, largest=True, sorted=True)  # :946
            score = best
        else:
            scale = seq_len.clamp(min=1).pow(-length_alpha).unsqueeze(2)         # :948
            ranked = (cand * scale).view(B, H * V)
            best, idx = torch.topk(ranked, k=H, dim=1, largest=True, sorted=True)  # :950
            score_normed = best
            score = flat.gather(1, idx)                                          # :951
        top_h1 = torch.topk(ranked, k=H + 1, dim=1, largest=True, sorted=True).values
        gaps = (top_h1[:, :-1] - top_h1[:, 1:]).nan_to_num(nan=float("inf"), posinf=float("inf"))
        margin = torch.minimum(margin, gaps.min(dim=1).values)
        parent = idx // V                                                        # :953
        new_tok = idx % V                                                        # :954
        gather_idx = parent.unsqueeze(2)
        if c > 1:
            tok[:, :, :c - 1] = tok[:, :, :c - 1].gather(1, gather_idx.expand(-1, -1, c - 1))  # :957
        tok[:, :, c - 1] = new_tok                                                # :958
        pad[:, :, :c] = pad[:, :, :c].gather(1, gather_idx.expand(-1, -1, c))     # :959
        if c < G:
            nxt_pad = (new_tok == 0) | pad[:, :, c - 1]                           # :963
            pad[:, :, c] = nxt_pad
            if early_exit and bool(nxt_pad.all()):                                # :964-967
                T = c
                break
            if guide_targets is not None:                                         # :969-971
                guide_mask = guide_mask.gather(1, gather_idx.expand(-1, -1, guide_mask.shape[2])) | \
                    (new_tok.unsqueeze(2) != guide_targets[:, c - 1].view(1, 1, -1))
            if vocab_mask is not None:                                            # :972-975
                vocab_mask = vocab_mask.gather(1, gather_idx.expand(-1, -1, vocab_mask.shape[2])) | \
                    (new_tok.unsqueeze(2) != vocab_targets[:, c - 1].view(1, 1, -1))
            if length_alpha != 0:
                seq_len = seq_len.gather(1, parent) + (~nxt_pad).to(dtype)        # :978
    tok = tok[:, :, :T].clone()
    pad = pad[:, :, :T].clone()
    tok.masked_fill_(pad, 0)                                                      # :980
    return dict(target=tok, padding=pad, score=score_normed if length_alpha != 0 else score, margin=margin)


# ----------------------------------------------------------------------------------------------------
# generate_all (embedding_decoder.py:986-1079): score every guide target by teacher forcing, keep the top-k
# ----------------------------------------------------------------------------------------------------

def generate_all(cfg: OracleCfg, sd: dict, embed: torch.Tensor, topk: int, temperature: float, length_alpha: float,
                 guide_targets: torch.Tensor, guide_renorm: bool, vocab_targets: Optional[torch.Tensor] = None,
                 vocab_per_token: bool = False, vocab_scaler: float = 0.0):
    """Returns dict(target B x K x C, padding B x K x C, score B x K sorted descending, all_scores B x W)."""
    W, Cmax = guide_targets.shape
    V = cfg.vocab_size
    pads = torch.zeros(W, Cmax, dtype=torch.bool)
    pads[:, 1:] = (guide_targets[:, :-1] == 0).cummax(dim=1).values                      # :991-994
    C = Cmax - int(pads.all(dim=0).sum())                                                 # :996
    pads = pads[:, :C]
    gt = guide_targets[:, :C].masked_fill(pads, 0)                                        # :997-998
    mism = torch.zeros(W, C, W, dtype=torch.bool)                                         # :1002-1005: target w' mismatches w before position c
    mism[:, 1:, :] = (gt[:, :-1, None] != gt.t()[None, :-1, :]).cummax(dim=1).values
    guide_scores = None
    if guide_renorm:                                                                      # :1006-1007
        idx = gt.t().expand(W, -1, -1).masked_fill(mism, V)
        guide_scores = torch.full((W, C, V + 1), NEG_INF).scatter_(2, idx, 0.0)[:, :, :-1]
    vocab_scores = None
    if vocab_targets is not None and vocab_scaler != 0:                                   # :1011-1036
        vt = vocab_targets[:, :C]
        Z = vt.shape[0]
        vm = torch.zeros(W, C, Z, dtype=torch.bool)
        vm[:, 1:, :] = (gt[:, :-1, None] != vt.t()[None, :-1, :]).cummax(dim=1).values
        vidx = vt.t().expand(W, -1, -1).masked_fill(vm, V)
        if vocab_per_token:
            vs = torch.zeros(W, C, V + 1).scatter_(2, vidx, 1.0)[:, :, :-1]
            vs = vs / vs.sum(dim=2, keepdim=True)
        else:
            cnt = torch.zeros(W, C, V + 1).scatter_add_(2, vidx, torch.ones(W, C, Z))
            vs = cnt[:, :, :-1] / (Z - cnt[:, :, -1:])
        vs = vs.gather(2, gt.unsqueeze(2)).squeeze(2).log().nan_to_num(nan=float("inf"), neginf=float("inf"), posinf=float("inf"))
        vocab_scores = vs.masked_fill(pads, 0.0).sum(dim=1) * vocab_scaler
    B = embed.shape[0]
    scores = torch.empty(B, W)
    for w0 in range(0, W, 64):
        tg = gt[w0:w0 + 64]
        M = tg.shape[0]
        full_t = tg.unsqueeze(0).expand(B, -1, -1).reshape(B * M, C)
        full_p = pads[w0:w0 + 64].unsqueeze(0).expand(B, -1, -1).reshape(B * M, C)
        logits, _ = forward_logits(cfg, sd, embed.repeat_interleave(M, dim=0), full_t, full_p, only_pred=False)   # :1065
        logits = logits.view(B, M, C, V) / temperature                                                             # :1066
        if guide_scores is not None:
            logits = logits + guide_scores[None, w0:w0 + 64]                                                        # :1068
        lp = torch.log_softmax(logits, dim=3).gather(3, tg[None, :, :, None].expand(B, -1, -1, 1)).squeeze(3)    # :1069-1070
        scores[:, w0:w0 + 64] = lp.masked_fill(pads[None, w0:w0 + 64], 0.0).sum(dim=2)                             # :1071-1072
    if vocab_scores is not None:
        scores = scores - vocab_scores.unsqueeze(0)                                        # :1075
    if length_alpha != 0:
        scores = scores * (C - pads.sum(dim=1)).clamp(min=1).float().pow(-length_alpha).unsqueeze(0)   # :1046-1050, :1077
    best, idx = torch.topk(scores, k=topk, dim=1, largest=True, sorted=True)             # :1079
    idx3 = idx.unsqueeze(2).expand(-1, -1, C)
    return dict(target=gt.expand(B, -1, -1).gather(1, idx3), padding=pads.expand(B, -1, -1).gather(1, idx3), score=best, all_scores=scores)


# ----------------------------------------------------------------------------------------------------
# Embedding noise (embedding_noise.py).  The random draws are explicit inputs so that the CUDA kernel's
# "debug" entry point (pre-drawn normals / uniforms) can be compared bit-for-bit-ish against this.
# ----------------------------------------------------------------------------------------------------

def _unit(x: torch.Tensor) -> torch.Tensor:
    return x / x.norm(dim=-1, keepdim=True).clamp_min(1e-12)  # F.normalize semantics


def noise_gauss_elem(embed: torch.Tensor, normals: torch.Tensor, vec_norm: float) -> torch.Tensor:
    """embedding_noise.py:72-75: e + (vec_norm/sqrt(F)) * n, renormalised."""
    sigma = vec_norm / math.sqrt(embed.shape[1])
    return _unit(embed + sigma * normals)


def noise_gauss_vec(embed: torch.Tensor, normals: torch.Tensor, row_normal: torch.Tensor, vec_norm: float) -> torch.Tensor:
    """embedding_noise.py:90-95: e + vec_norm * g * unit(n) with one scalar g ~ N(0,1) per row."""
    return _unit(embed + vec_norm * row_normal.view(-1, 1) * _unit(normals))


def noise_angle(embed: torch.Tensor, normals: torch.Tensor, angle_rad: torch.Tensor) -> torch.Tensor:
    """embedding_noise.py:105-112: rotate e by `angle` towards the direction of n orthogonal to e."""
    d = _unit(normals - embed * (embed * normals).sum(dim=1, keepdim=True))
    a = angle_rad.view(-1, 1)
    return _unit(embed * torch.cos(a) + d * torch.sin(a))


def gauss_angle(row_normal: torch.Tensor, angle_std_deg: float, angle_max_deg: float) -> torch.Tensor:
    """embedding_noise.py:131-132."""
    m = math.radians(angle_max_deg)
    return (row_normal * math.radians(angle_std_deg)).clamp(min=-m, max=m)


def uniform_angle(row_uniform: torch.Tensor, angle_min_deg: float, angle_max_deg: float) -> torch.Tensor:
    """embedding_noise.py:151-152: U[min, max) from a U[0,1) draw."""
    lo, hi = math.radians(angle_min_deg), math.radians(angle_max_deg)
    return lo + (hi - lo) * row_uniform


def noise_gauss_elem_uniform_angle(embed: torch.Tensor, normals_angle: torch.Tensor, angle_uniform: torch.Tensor,
                                   normals_elem: torch.Tensor, mix_uniform: torch.Tensor, vec_norm: float,
                                   angle_min_deg: float, angle_max_deg: float, mix_ratio: float) -> torch.Tensor:
    """embedding_noise.py:169-172: per row, angle noise with probability mix_ratio, else element noise."""
    a = noise_angle(embed, normals_angle, uniform_angle(angle_uniform, angle_min_deg, angle_max_deg))
    g = noise_gauss_elem(embed, normals_elem, vec_norm)
    return torch.where((mix_uniform < mix_ratio).view(-1, 1), a, g)


# ----------------------------------------------------------------------------------------------------
# Synthetic-weight helper shared by tests (the product has its own copy in novic_b200/synth.py; the two
# are compared in tests so neither silently drifts).
# ----------------------------------------------------------------------------------------------------

def reference_init_stds(cfg: OracleCfg) -> dict:
    """Standard deviations the reference's 'balanced' init produces for the default architecture
    (embedding_decoder.py:203-226, :228-278, :329-407; utils.py:84-112)."""
    E, L, K, P = cfg.hidden_dim, cfg.num_layers, cfg.ffn_dim, cfg.prefix_len
    f = 1.0 / math.sqrt(E)
    lf = 1.0 / math.sqrt(2 * L)
    attn_scale = math.sqrt((1 + (P - 1) / P) / P)
    return {
        "embed_mlp": 1.0 / math.sqrt(2.0),
        "logits": 1.0 / math.sqrt(2.0),
        "pos": 1.0 / math.sqrt(2.0),
        "in_proj": f,
        "out_proj": f / attn_scale * lf,
        "linear1": f,
        "linear2": 1.0 / (math.sqrt(K) * 0.6521) * lf,
        "norm": 1.0,
        "final_norm": f,
    }
