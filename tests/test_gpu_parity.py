"""Parity of the CUDA path (through the C ABI) against the reference's committed outputs and the CPU oracle.

Tolerance (BASELINE.json north_star / SURVEY.md section 8c): the kernels compute with bf16 operands and fp32
accumulation, the reference in fp32, so
    |logit_cuda - logit_ref| <= LOGIT_TOL = 0.06
and token ids / correctness flags must be identical wherever the reference's top-2 logit margin exceeds
MARGIN_TOL = 0.12 (= 2 * LOGIT_TOL).  Integer outputs (padding, lengths, loss basis, T) are exact.
"""
import numpy as np
import pytest
import torch

from novic_b200 import _abi, default_decoder, synth
from oracle import novic_oracle as orc
from tests.golden_util import B_GOLD, Golden, gold_embed, weight_case

pytestmark = pytest.mark.gpu

LOGIT_TOL = 0.06
MARGIN_TOL = 0.12
DEV = "cuda:0"
TAGS = ("lively", "eos", "eosall")


@pytest.fixture(scope="module")
def gold():
    return Golden()


@pytest.fixture(scope="module")
def models():
    cache = {}

    def get(tag):
        if tag not in cache:
            cache[tag] = default_decoder(synth.DecoderDims(), weight_case(tag)).to(DEV)
        return cache[tag]
    return get


def test_library_is_the_cuda_path(built_lib):
    assert built_lib.novic_version() >= 1
    before = built_lib.novic_launch_count()
    m = default_decoder(synth.DecoderDims(), weight_case("lively")).to(DEV)
    with torch.inference_mode():
        m.generate(gold_embed().to(DEV), False, True, 1.0, 0.0, None, None, False)
    assert built_lib.novic_launch_count() - before > 100  # kernels were really launched through the library


@pytest.mark.parametrize("block_n", [128, 256, 512])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 128, 512), (256, 384, 512), (100, 200, 128), (1, 6912, 512), (4096, 1536, 512), (777, 2048, 1024)])
def test_tcgen05_gemm_building_block(built_lib, M, N, K, block_n):
    """The persistent tcgen05 GEMM against an fp64 matmul of the same bf16 operands: 128 x 128 tiles (block_n 128), 128 x 256 tiles (256) and
    256 x 256 tiles computed by CTA pairs with tcgen05.mma.cta_group::2 (512; ragged M / N exercise the pair whose second CTA has no rows)."""
    if block_n != 128 and K % 128 != 0:
        pytest.skip("the wide-stage kernels take two k-blocks per request")
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    a = (torch.randn(M, K, generator=g) * 0.5).bfloat16().to(DEV)
    w = (torch.randn(N, K, generator=g) * 0.5).bfloat16().to(DEV)
    out = torch.full((M, N), float("nan"), device=DEV)
    _abi.check(built_lib.novic_debug_gemm(a.data_ptr(), w.data_ptr(), out.data_ptr(), M, N, K, block_n, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = a.double() @ w.double().t()
    assert (out.double() - ref).abs().max().item() < 1e-3


@pytest.mark.parametrize("tag", TAGS)
def test_teacher_forced_forward_vs_reference_outputs(gold, models, tag):
    dims = synth.DecoderDims()
    tgt, pad = synth.synth_targets(B_GOLD, dims, seed=5)
    with torch.inference_mode():
        logits, epad, ls, lb, cor = models(tag)(gold_embed().to(DEV), tgt.to(DEV), pad.to(DEV), None, True, True, False, None)
    logits, epad, cor = logits.cpu(), epad.cpu(), cor.cpu()
    valid = ~pad
    assert torch.equal(epad, gold[f"{tag}/tf/effpad"])
    assert (logits[..., gold["meta/probes"]] - gold[f"{tag}/tf/probes"])[valid].abs().max() <= LOGIT_TOL
    assert (torch.logsumexp(logits, -1) - gold[f"{tag}/tf/lse"])[valid].abs().max() <= LOGIT_TOL
    assert (logits.gather(-1, tgt.unsqueeze(-1)).squeeze(-1) - gold[f"{tag}/tf/at_target"])[valid].abs().max() <= LOGIT_TOL
    decided = valid & (gold[f"{tag}/tf/margin"] > MARGIN_TOL)
    assert decided.float().mean() > 0.02
    assert torch.equal(logits.argmax(-1)[decided], gold[f"{tag}/tf/argmax"][decided])
    assert torch.equal(cor[decided], gold[f"{tag}/tf/correct"][decided])
    assert not cor[pad].any()
    n = int(gold[f"{tag}/tf/loss"][1].item())
    assert int(lb) == n
    assert abs(ls.item() - gold[f"{tag}/tf/loss"][0].item()) <= 2 * LOGIT_TOL * n


@pytest.mark.parametrize("tag", ("lively", "eos"))
def test_multi_target_weighted_forward_vs_reference_outputs(gold, models, tag):
    dims = synth.DecoderDims()
    tgt3, pad3 = synth.synth_targets(8, dims, seed=6, multi=3)
    w3 = torch.from_numpy(np.random.default_rng(8).random((8, 3)).astype(np.float32))
    w3[1, 2] = 0.0
    with torch.inference_mode():
        logits, epad, ls, lb, cor = models(tag)(gold_embed()[:8].to(DEV), tgt3.to(DEV), pad3.to(DEV), w3.to(DEV), True, True, False, None)
    assert logits.shape == (8, 3, 16, dims.vocab_size) and epad.shape == (8, 3, 16) and cor.shape == (8, 3, 16)
    assert torch.equal(epad.cpu(), gold[f"{tag}/tfm/effpad"])
    valid = ~gold[f"{tag}/tfm/effpad"]
    assert (logits.cpu()[..., gold["meta/probes"]] - gold[f"{tag}/tfm/probes"])[valid].abs().max() <= LOGIT_TOL
    assert abs(lb.item() - gold[f"{tag}/tfm/loss"][1].item()) < 1e-3
    assert abs(ls.item() - gold[f"{tag}/tfm/loss"][0].item()) <= 2 * LOGIT_TOL * gold[f"{tag}/tfm/loss"][1].item()


def _first_divergence(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Per row: index of the first differing column (or ncols if none)."""
    n = min(a.shape[1], b.shape[1])
    diff = a[:, :n] != b[:, :n]
    first = torch.where(diff.any(dim=1), diff.float().argmax(dim=1), torch.full((a.shape[0],), n))
    return first


@pytest.mark.parametrize("tag", TAGS)
@pytest.mark.parametrize("name,tau,alpha", (("g10", 1.0, 0.0), ("g07", 0.7, 0.5)))
def test_greedy_vs_reference_outputs(gold, models, tag, name, tau, alpha):
    with torch.inference_mode():
        tok, pad, lg, ls, lb, sc = models(tag).generate(gold_embed().to(DEV), True, True, tau, alpha, None, None, False)
    tok, pad, lg, sc = tok.cpu(), pad.cpu(), lg.cpu(), sc.cpu()
    g_tok, g_pad, g_margin = gold[f"{tag}/{name}/tok"], gold[f"{tag}/{name}/pad"], gold[f"{tag}/{name}/margin"]
    first = _first_divergence(tok, g_tok)
    T = g_tok.shape[1]
    clean = first >= min(T, tok.shape[1])
    # a row may only leave the reference's trajectory at a step where the reference itself was undecided
    for b in (~clean).nonzero().flatten().tolist():
        assert g_margin[b, first[b]] <= MARGIN_TOL, f"row {b} diverged at step {first[b]} with margin {g_margin[b, first[b]]:.3f}"
    assert clean.float().mean() >= 0.9
    if clean.all():
        assert tok.shape == g_tok.shape and torch.equal(pad, g_pad)  # same early-exit length T, same padding
        assert int(lb) == int(gold[f"{tag}/{name}/loss"][1].item())
        assert abs(ls.item() - gold[f"{tag}/{name}/loss"][0].item()) <= 2 * LOGIT_TOL * int(lb)
    n = min(T, tok.shape[1])
    assert torch.equal(pad[clean, :n], g_pad[clean, :n])
    n_tok = (~g_pad).sum(dim=1).float()
    assert ((sc - gold[f"{tag}/{name}/score"]).abs()[clean] <= 2 * LOGIT_TOL / tau * n_tok[clean].clamp(min=1).pow(1 - alpha) + 1e-3).all()
    live = (~g_pad[:, :n]) & clean.unsqueeze(1)
    assert (lg[:, :n][..., gold["meta/probes"]] - gold[f"{tag}/{name}/probes"][:, :n])[live].abs().max() <= LOGIT_TOL
    assert (torch.logsumexp(lg[:, :n], -1) - gold[f"{tag}/{name}/lse"][:, :n])[live].abs().max() <= LOGIT_TOL
    assert (tok[pad] == 0).all()


def _oracle_sequence_scores(cfg, sd, embed, tok, pad, tau, alpha):
    """Score arbitrary candidate sequences the way beam search does: sum of log softmax(logits / tau) over the unpadded
    tokens (the first generated token is never the end token, but that only removes candidates, not probability mass)."""
    B, H, T = tok.shape
    full = torch.zeros(B * H, cfg.token_length, dtype=torch.int64)
    full[:, :T] = tok.reshape(B * H, T)
    fpad = torch.ones(B * H, cfg.token_length, dtype=torch.bool)
    fpad[:, :T] = pad.reshape(B * H, T)
    logits, _ = orc.forward_logits(cfg, sd, embed, full, fpad, only_pred=False)
    lp = torch.log_softmax(logits / tau, dim=-1).gather(-1, full.unsqueeze(-1)).squeeze(-1).masked_fill(fpad, 0.0)
    s = lp.sum(dim=1)
    if alpha != 0:
        s = s * (~fpad).sum(dim=1).clamp(min=1).float().pow(-alpha)
    return s.view(B, H)


@pytest.mark.parametrize("tag", TAGS)
@pytest.mark.parametrize("name,H,tau,alpha", (("b3", 3, 1.0, 0.0), ("b5", 5, 1.3, 0.6), ("b10", 10, 1.0, 0.0)))
def test_beam_vs_reference_outputs(gold, models, tag, name, H, tau, alpha):
    sd = weight_case(tag)
    cfg = orc.cfg_from_state_dict(sd)
    with torch.inference_mode():
        tok, pad, sc = models(tag).generate_beam(gold_embed().to(DEV), H, tau, alpha, None, False, 0.0, None, False)
        tok, pad, sc = tok.cpu(), pad.cpu(), sc.cpu()
        g_tok, g_pad, g_sc = gold[f"{tag}/{name}/tok"], gold[f"{tag}/{name}/pad"], gold[f"{tag}/{name}/score"]
        assert tok.shape[:2] == g_tok.shape[:2] and abs(tok.shape[2] - g_tok.shape[2]) <= 1
        assert (sc[:, :-1] >= sc[:, 1:]).all()                      # sorted descending like torch.topk(sorted=True)
        assert (tok[pad] == 0).all()
        n_steps = tok.shape[2]
        tol = 2 * LOGIT_TOL / tau * n_steps
        # (1) the scores the kernel reports are the oracle's scores of the sequences it returned
        rescored = _oracle_sequence_scores(cfg, sd, gold_embed(), tok, pad, tau, alpha)
        assert (rescored - sc).abs().max() <= tol
        # (2) Beam search prunes greedily over H * V candidates whose neighbouring ranks are always closer than any bf16
        #     tolerance (the reference search's own pruning gaps are ~1e-3, see `margin` in oracle.generate_beam), so
        #     literal equality cannot be demanded row by row; what must hold is:
        #     - beams that are literally the reference's carry the reference's score and padding,
        #     - the large majority of beams (and of best beams) are literally identical,
        #     - search quality is the same: on (almost) every sample the best beam scores like the reference's best beam.
        if tok.shape == g_tok.shape:
            same = (tok == g_tok).all(dim=2)
            assert same[:, 0].float().mean() >= 0.8
            assert same.float().mean() >= 0.6
            assert torch.equal(pad[same], g_pad[same])
            assert (sc - g_sc)[same].abs().max() <= tol
        assert ((sc[:, 0] - g_sc[:, 0]).abs() <= tol).float().mean() >= 0.9


@pytest.mark.parametrize("B,C,use_pad,only_pred", [(1, 1, False, False), (3, 2, True, False), (129, 16, True, False), (40, 7, False, True), (64, 16, True, True)])
def test_forward_shapes_vs_oracle(models, B, C, use_pad, only_pred):
    dims = synth.DecoderDims()
    sd = weight_case("lively")
    cfg = orc.cfg_from_state_dict(sd)
    embed = synth.synth_embeddings(B, seed=77)
    tgt, pad = synth.synth_targets(B, dims, seed=B + C)
    tgt, pad = tgt[:, :C].contiguous(), (pad[:, :C].contiguous() if use_pad else None)
    with torch.inference_mode():
        o_logits, o_pad = orc.forward_logits(cfg, sd, embed, tgt, pad, only_pred)
        logits, epad, _, _, _ = models("lively")(embed.to(DEV), tgt.to(DEV), None if pad is None else pad.to(DEV), None, False, False, only_pred, None)
    assert logits.shape == o_logits.shape
    if pad is None:
        assert epad is None
        valid = torch.ones(o_logits.shape[:2], dtype=torch.bool)
    else:
        assert torch.equal(epad.cpu(), o_pad)
        valid = ~o_pad
    assert (logits.cpu() - o_logits)[valid].abs().max() <= LOGIT_TOL


def test_greedy_ragged_batch_sizes_and_chunking(models, monkeypatch):
    """Edge cases: B = 1, B not a multiple of any tile, and the chunked path (B > MAX_SEQS_PER_CALL) must agree with the
    single-call path bit for bit."""
    from novic_b200 import decoder as dec_mod
    model = models("eos")
    embed = synth.synth_embeddings(150, seed=5).to(DEV)
    with torch.inference_mode():
        full = model.generate(embed, False, True, 1.0, 0.0, None, None, False)
        one = model.generate(embed[:1], False, True, 1.0, 0.0, None, None, False)
        monkeypatch.setattr(dec_mod, "MAX_SEQS_PER_CALL", 64)
        chunked = model.generate(embed, False, True, 1.0, 0.0, None, None, False)
        beam_chunked = model.generate_beam(embed, 3, 1.0, 0.0, None, False, 0.0, None, False)
        monkeypatch.setattr(dec_mod, "MAX_SEQS_PER_CALL", 1 << 15)
        beam_full = model.generate_beam(embed, 3, 1.0, 0.0, None, False, 0.0, None, False)
    n = min(one[0].shape[1], full[0].shape[1])
    assert torch.equal(one[0][:, :n], full[0][:1, :n])
    assert torch.equal(chunked[0], full[0]) and torch.equal(chunked[1], full[1]) and torch.equal(chunked[5], full[5])
    assert chunked[4].item() == full[4].item()
    assert torch.equal(beam_chunked[0], beam_full[0]) and torch.equal(beam_chunked[2], beam_full[2])


def test_full_size_greedy_properties(models):
    """BASELINE config #2 (B = 4096, random-init default decoder): size-independent properties.
    - random-init weights never emit the end token: T = G = 15, no padding
    - determinism: two runs are bit-identical
    - the KV-cached decode loop is consistent with the teacher-forced forward: feeding the generated ids back through
      forward() reproduces them as arg-max (wherever decided) and reproduces the scores (sum of log-probs)."""
    dims = synth.DecoderDims()
    model = default_decoder(dims, synth.synth_state_dict(dims, seed=1)).to(DEV)
    embed = synth.synth_embeddings(4096, seed=1234).to(DEV)
    with torch.inference_mode():
        tok, pad, _, ls, lb, sc = model.generate(embed, False, True, 1.0, 0.0, None, None, False)
        tok2, _, _, _, _, sc2 = model.generate(embed, False, True, 1.0, 0.0, None, None, False)
        assert tok.shape == (4096, 15) and not pad.any() and int(lb) == 4096 * 15
        assert torch.equal(tok, tok2) and torch.equal(sc, sc2)
        assert (tok > 0).all() and (tok < dims.vocab_size).all()
        sub = slice(0, 512)
        full = torch.cat((tok[sub], torch.zeros(512, 1, dtype=torch.int64, device=DEV)), dim=1)
        logits, _, _, _, _ = model(embed[sub], full, None, None, False, False, False, None)
        lp = torch.log_softmax(logits[:, :15], dim=-1)
        top2 = logits[:, :15].topk(2, dim=-1).values
        decided = (top2[..., 0] - top2[..., 1]) > MARGIN_TOL
        first_ok = torch.ones_like(decided)
        am = logits[:, :15].argmax(-1)
        am[:, 0] = logits[:, 0, 1:].argmax(-1) + 1
        assert torch.equal(am[decided & first_ok], tok[sub][decided & first_ok])
        assert (lp.gather(-1, tok[sub].unsqueeze(-1)).squeeze(-1).sum(dim=1) - sc[sub]).abs().max() <= LOGIT_TOL * 15
        assert abs(-sc.sum().item() - ls.item()) <= 1e-3 * abs(ls.item())  # tau = 1, alpha = 0: loss_sum = -sum(score)


def test_collect_logits_flag_and_calc_loss_flag(models):
    model = models("lively")
    embed = gold_embed().to(DEV)
    with torch.inference_mode():
        a = model.generate(embed, False, False, 1.0, 0.0, None, None, False)
        b = model.generate(embed, True, True, 1.0, 0.0, None, None, False)
    assert a[2] is None and a[3] is None and a[4] is None and a[5] is None
    assert b[2].shape == (B_GOLD, b[0].shape[1], synth.DecoderDims().vocab_size)
    assert torch.equal(a[0], b[0])
    w = torch.rand(B_GOLD, device=DEV)
    with torch.inference_mode():
        c = model.generate(embed, False, True, 1.0, 0.0, w, None, False)
    assert abs(c[4].item() - (w * (~c[1]).sum(dim=1)).sum().item()) < 1e-2


def test_noise_vs_reference_outputs(gold):
    from novic_b200 import noise
    e0 = synth.synth_embeddings(16, seed=9)
    g = lambda k: gold[f"noise/{k}"].to(DEV)
    cases = [
        (noise.GaussElemNoise(1024, 3.25), ("gauss_elem/na", None, None, None), "gauss_elem/out"),
        (noise.GaussVecNoise(1024, 0.8), ("gauss_vec/na", None, "gauss_vec/ra", None), "gauss_vec/out"),
        (noise.UniformAngleNoise(1024, 45.0, 75.0), ("uniform_angle/na", None, "uniform_angle/ra", None), "uniform_angle/out"),
        (noise.GaussAngleNoise(1024, 30.0, 40.0), ("gauss_angle/na", None, "gauss_angle/ra", None), "gauss_angle/out"),
        (noise.GaussElemUniformAngleNoise(1024, 3.25, 45.0, 75.0, 0.5), ("mix/na", "mix/nb", "mix/ra", "mix/rb"), "mix/out"),
    ]
    for mod, keys, out_key in cases:
        na, nb, ra, rb = [None if k is None else g(k).contiguous() for k in keys]
        got = mod.apply_predrawn(e0.clone().to(DEV), na, nb, ra, rb)
        assert (got - g(out_key)).abs().max().item() < 2e-6  # fp32 arithmetic: same formulae, different summation order


def test_noise_statistics_and_in_place_contract():
    import math
    from novic_b200 import EmbeddingNoise
    e = synth.synth_embeddings(8192, seed=3).to(DEV)
    torch.manual_seed(0)
    for scheme, check in (
        ("GaussElem", lambda c, a: abs(c.mean().item() - 1 / math.sqrt(1 + 3.25 ** 2)) < 0.01),
        ("UniformAngle", lambda c, a: a.min() >= 45.0 - 1e-2 and a.max() <= 75.0 + 1e-2 and abs(a.mean().item() - 60.0) < 0.5),
        ("GaussAngle", lambda c, a: a.max() <= 75.0 + 1e-2 and abs(a.std().item() - 30.0 * 0.6028) < 1.5),  # |N(0, 30)| clamped at 75
        ("GaussVec", lambda c, a: c.min() > 0.0),
        ("GaussElemUniformAngle", lambda c, a: 0.28 < c.mean().item() < 0.36 and a.min() >= 45.0 - 1e-2),
    ):
        mod = EmbeddingNoise.create(scheme, 1024, 3.25, 45.0, 75.0, 30.0, 0.15)
        x = e.clone()
        y = mod(x)
        assert y.data_ptr() == x.data_ptr()                              # in place (embedding_noise.py:50)
        assert (y.norm(dim=1) - 1).abs().max().item() < 1e-5             # unit-norm preserving
        cos = (y * e).sum(dim=1).clamp(-1, 1)
        assert check(cos, torch.rad2deg(torch.acos(cos))), scheme
        z = mod(e.clone())
        assert not torch.equal(y, z)                                     # the stream advances between calls
    assert EmbeddingNoise.create("", 1024, 1, 0, 0, 0, 0) is None
    with pytest.raises(ValueError):
        EmbeddingNoise.create("nope", 1024, 1, 0, 0, 0, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ("lively", "eos", "eosall"))
def test_guided_correctness_evaluation_vs_reference_outputs(gold, models, tag):
    """forward(..., guide_targets) in evaluation mode (embedding_decoder.py:754-760): `correct` must equal the reference's wherever the
    guided top-2 margin exceeds the logit tolerance; logits / loss are those of the unguided forward."""
    from tests.golden_util import guided_eval_case
    gt, tgt, pad = guided_eval_case()
    m = models(tag)
    with torch.inference_mode():
        lg, opad, ls, lb, cor = m(gold_embed().to(DEV), tgt.to(DEV), pad.to(DEV), None, True, True, False, gt.to(DEV))
        lg0, _, ls0, lb0, cor0 = m(gold_embed().to(DEV), tgt.to(DEV), pad.to(DEV), None, True, True, False, None)
    assert torch.equal(lg, lg0) and ls.item() == ls0.item() and int(lb) == int(lb0)
    ref_cor, margin = gold[f"{tag}/tfg/correct"], gold[f"{tag}/tfg/margin"]
    decided = (margin > 0.12) & ~pad
    assert torch.equal(cor.cpu()[decided], ref_cor[decided])
    assert not cor.cpu()[pad].any() and cor.sum() > cor0.sum()
    with pytest.raises(AssertionError):
        m(gold_embed().to(DEV), tgt.to(DEV), pad.to(DEV), None, True, True, True, gt.to(DEV))      # only_pred is incompatible (embedding_decoder.py:755)


@pytest.mark.gpu
def test_forward_without_targets_is_the_first_step(models):
    """forward(embed, target=None, ...) returns the logits of the first generated position (B x 1 x V), i.e. step 1 of generate."""
    m = models("lively")
    sd = weight_case("lively")
    cfg = orc.cfg_from_state_dict(sd)
    e = gold_embed()[:12]
    with torch.inference_mode():
        lg, pad, ls, lb, cor = m(e.to(DEV), None, None, None, False, False, False, None)
        o, _ = orc.forward_logits(cfg, sd, e, None, None, only_pred=False)
        g = m.generate(e.to(DEV), True, False, 1.0, 0.0, None, None, False)
    assert lg.shape == (12, 1, cfg.vocab_size) and pad is None and ls is None and lb is None and cor is None
    assert (lg.cpu() - o).abs().max().item() <= 0.06
    assert (lg[:, 0] - g[2][:, 0]).abs().max().item() <= 1e-3          # same kernels, prefill path vs decode path


@pytest.mark.gpu
def test_qkv_tail_of_the_block_kernel_is_bit_identical(monkeypatch):
    """The next layer's QKV projection in the tail of the fused block kernel (opt-in, NOVIC_FUSE_QKV=1: measured slower than the
    stand-alone launches) against the stand-alone QKV GEMM launches (default): same operands, same accumulation
    order -> bit-identical ids, scores, logits, for greedy (ragged batch: the last 128-row block is partial), beam search and the
    teacher-forced forward.  The small graph-replayed batches are the regression case of the decode step's launch fence: with a
    few CTAs per kernel a whole step's prologues fit on the GPU at once, and the attention kernel's early K/V requests read rows of
    the prefix pass before it had finished until select_greedy_kernel was made to wait before releasing its dependents."""
    dims = synth.DecoderDims()
    sd = weight_case("eos")
    embed = synth.synth_embeddings(300, seed=5).to(DEV)
    tgt, pad = synth.synth_targets(40, dims, seed=3)
    outs = []
    monkeypatch.setenv("NOVIC_FFN1_KSPLIT", "0")     # the tail lives in the block kernel that gathers the LN2 rows
    monkeypatch.setenv("NOVIC_BLOCK_ROWS", "128")    # ... on 128-row tiles (the default row-owner kernel sums in another order)
    for flag in ("0", "1"):                      # the switch is read when a handle is created
        monkeypatch.setenv("NOVIC_FUSE_QKV", flag)
        m = default_decoder(dims, sd).to(DEV)
        with torch.inference_mode():
            g = m.generate(embed, True, True, 0.9, 0.2, None, None, False)
            b = m.generate_beam(embed[:64], 3, 1.0, 0.0, None, False, 0.0, None, False)
            f = m(embed[:40], tgt.to(DEV), pad.to(DEV), None, True, True, False, None)
            small = [m.generate(embed[:n], False, True, 1.0, 0.0, None, None, False) for n in (1, 2, 5, 31, 33) for _ in range(2)]
        outs.append((g, b, f, small))
        del m
    (g0, b0, f0, s0), (g1, b1, f1, s1) = outs
    for a, c in zip(s0, s1):
        assert torch.equal(a[0], c[0]) and torch.equal(a[1], c[1]) and torch.equal(a[5], c[5])
    assert torch.equal(g0[0], g1[0]) and torch.equal(g0[1], g1[1]) and torch.equal(g0[2], g1[2]) and torch.equal(g0[5], g1[5])
    assert torch.equal(b0[0], b1[0]) and torch.equal(b0[2], b1[2])
    assert torch.equal(f0[0], f1[0]) and f0[2].item() == f1[2].item() and torch.equal(f0[4], f1[4])


def test_block_kernel_on_64_row_tiles_is_bit_identical(monkeypatch):
    """The fused out-proj + feed-forward block kernel on 64-row tiles (outproj_ffn64_kernel: what small row counts use, NOVIC_BLOCK_ROWS=64
    forces it everywhere) against the 128-row kernel (NOVIC_BLOCK_ROWS=128): same operands and accumulation order -> bit-identical ids,
    scores and logits for greedy (ragged: 300 = 4 x 64 + 44 rows per decode step, 1200 prefix rows), beam search and the
    teacher-forced forward; and with two CTAs resident per SM (NOVIC_BLOCK64_PAD=0) as well."""
    dims = synth.DecoderDims()
    sd = weight_case("eos")
    embed = synth.synth_embeddings(300, seed=5).to(DEV)
    tgt, pad = synth.synth_targets(40, dims, seed=3)
    outs = []
    monkeypatch.setenv("NOVIC_FFN1_KSPLIT", "0")     # the 128-row kernel that gathers the LN2 rows, as the 64-row one does
    for rows, extra in (("128", "16384"), ("64", "16384"), ("64", "0")):     # the switches are read when a handle is created
        monkeypatch.setenv("NOVIC_BLOCK_ROWS", rows)
        monkeypatch.setenv("NOVIC_BLOCK64_PAD", extra)
        m = default_decoder(dims, sd).to(DEV)
        with torch.inference_mode():
            g = m.generate(embed, True, True, 0.9, 0.2, None, None, False)
            b = m.generate_beam(embed[:64], 3, 1.0, 0.0, None, False, 0.0, None, False)
            f = m(embed[:40], tgt.to(DEV), pad.to(DEV), None, True, True, False, None)
            small = [m.generate(embed[:n], False, True, 1.0, 0.0, None, None, False) for n in (1, 63, 65)]
        outs.append((g, b, f, small))
        del m
    g0, b0, f0, s0 = outs[0]
    for g1, b1, f1, s1 in outs[1:]:
        assert torch.equal(g0[0], g1[0]) and torch.equal(g0[1], g1[1]) and torch.equal(g0[2], g1[2]) and torch.equal(g0[5], g1[5])
        assert torch.equal(b0[0], b1[0]) and torch.equal(b0[2], b1[2])
        assert torch.equal(f0[0], f1[0]) and f0[2].item() == f1[2].item() and torch.equal(f0[4], f1[4])
        for a, c in zip(s0, s1):
            assert torch.equal(a[0], c[0]) and torch.equal(a[1], c[1]) and torch.equal(a[5], c[5])


def test_block_kernel_with_k_split_ffn1_agrees_to_rounding(monkeypatch):
    """The default 128-row block kernel (outproj_ffn_ks_kernel: every CTA multiplies its own 128 LN2 columns and the partial FFN1 sums are
    reduce-scattered over the cluster) against the variant that gathers the whole LN2 row first (NOVIC_FFN1_KSPLIT=0).  The hidden
    pre-activations are sums of four K = 128 partial products instead of one K = 512 accumulation, so the two agree to fp32 rounding before
    the bf16 casts and to a few bf16 flips after six layers: teacher-forced logits within 0.01 (measured: max 3.4e-3, mean 2e-4 at a
    logit rms of 0.18; the tolerance against the reference is 0.06), greedy ids identical on >= 98 % of the rows (a row may leave
    the other variant's trajectory only at a near-tie)."""
    dims = synth.DecoderDims()
    sd = weight_case("lively")
    embed = synth.synth_embeddings(2048, seed=5).to(DEV)       # 2048 rows per decode step: above the 64-row kernel's range
    tgt, pad = synth.synth_targets(160, dims, seed=3)
    outs = []
    for flag in ("0", "1"):
        monkeypatch.setenv("NOVIC_FFN1_KSPLIT", flag)
        monkeypatch.setenv("NOVIC_BLOCK_ROWS", "128")
        m = default_decoder(dims, sd).to(DEV)
        with torch.inference_mode():
            f = m(embed[:160], tgt.to(DEV), pad.to(DEV), None, True, True, False, None)
            g = m.generate(embed, False, True, 1.0, 0.0, None, None, False)
        outs.append((f[0].float().cpu(), g[0].cpu()))
        del m
    (f0, g0), (f1, g1) = outs
    valid = ~pad
    assert (f0 - f1)[valid].abs().max().item() <= 0.01
    n = min(g0.shape[1], g1.shape[1])
    same = (g0[:, :n] == g1[:, :n]).all(dim=1).float().mean().item()
    print(f"rows with identical greedy ids: {same:.4f}")
    assert same >= 0.98


def test_row_owner_block_kernel_agrees_with_the_cluster_kernel_to_rounding(monkeypatch):
    """The default block kernel (block_rows_kernel: one CTA per 32 rows, the weights streamed as the M operand of the MMAs, LayerNorm
    statistics summed over TMEM lanes instead of over a thread's registers) against the 128-row cluster kernel that gathers the whole
    LN2 row (NOVIC_BLOCK_ROWS=128, NOVIC_FFN1_KSPLIT=0).  Same products, K = 512 accumulated in one chain in both; the statistics are
    summed in a different order, so the two agree to fp32 rounding before the bf16 casts and to a few bf16 flips after six layers:
    teacher-forced logits within 0.01 (tolerance against the reference: 0.06), greedy ids identical on >= 98 % of the rows.  Ragged row
    counts: 2049 rows per decode step (64 full 32-row blocks + one row), 8196 prefix rows, 160 x 16 teacher-forced rows with the row
    remap of the last layer, and single-block passes of 1 / 31 / 33 rows.  The third leg runs the same kernel on 64 rows per CTA
    (block_rows64_kernel, what passes of more than 4736 rows use: the residual rows live in TMEM and the out-proj / FFN2 MMAs accumulate
    onto them; NOVIC_BLOCK_ROWS64_MIN=1 forces it for every pass - a last CTA with one row, a last 32-row block that does not exist)."""
    dims = synth.DecoderDims()
    sd = weight_case("lively")
    embed = synth.synth_embeddings(2049, seed=5).to(DEV)
    tgt, pad = synth.synth_targets(160, dims, seed=3)
    outs = []
    for rows, min64 in (("128", "0"), ("32", "0"), ("32", "1")):     # cluster kernel; 32 rows per CTA everywhere; 64 rows per CTA everywhere
        monkeypatch.setenv("NOVIC_FFN1_KSPLIT", "0")
        monkeypatch.setenv("NOVIC_BLOCK_ROWS", rows)
        monkeypatch.setenv("NOVIC_BLOCK_ROWS64_MIN", min64)
        m = default_decoder(dims, sd).to(DEV)
        with torch.inference_mode():
            f = m(embed[:160], tgt.to(DEV), pad.to(DEV), None, True, True, False, None)
            g = m.generate(embed, False, True, 1.0, 0.0, None, None, False)
            small = [m.generate(embed[:n], True, True, 1.0, 0.0, None, None, False) for n in (1, 31, 33)]
        outs.append((f[0].float().cpu(), g[0].cpu(), [(s[0].cpu(), s[2].float().cpu()) for s in small]))
        del m
    f0, g0, s0 = outs[0]
    for (f1, g1, s1), floor in zip(outs[1:], (0.98, 0.965)):
        d = (f0 - f1)[~pad].abs()
        assert d.max().item() <= 0.01
        n = min(g0.shape[1], g1.shape[1])
        same = (g0[:, :n] == g1[:, :n]).all(dim=1).float().mean().item()
        print(f"teacher-forced logits: max |d| {d.max().item():.2e}, mean {d.mean().item():.2e}; rows with identical greedy ids: {same:.4f}")
        # measured: 0.9966 (32 rows per CTA), 0.9800 (64 rows: the residual is also summed in another order - inside the tensor core)
        assert same >= floor
        for (t0, l0), (t1, l1) in zip(s0, s1):
            assert (l0[:, 0] - l1[:, 0]).abs().max().item() <= 0.01          # first step: same inputs, no trajectory effects
            assert torch.equal(t0[:, 0], t1[:, 0]) or (l0[:, 0].topk(2, dim=-1).values.diff(dim=-1).abs().min().item() <= 0.02)


def test_attention_inside_the_block_kernel_is_bit_identical(monkeypatch):
    """NOVIC_FUSE_ATTN=1: the decode-step attention of a CTA's 32 sequences runs on the epilogue warps of the row-owner block kernel
    (block_rows_kernel<true>, K / V staged in two slots of the weight ring, the attention row written straight into the resident MMA
    operand) instead of as a launch of its own.  Same arithmetic in the same order -> bit-identical ids, padding, scores and step
    logits for greedy (ragged: 301 rows = 9 blocks + 13 rows, so some warps own one sequence and some none), beam search (ancestry
    tables, per-row K / V requests) and guided decoding keeps working on the unfused kernels of its masked passes."""
    dims = synth.DecoderDims()
    sd = weight_case("eos")
    embed = synth.synth_embeddings(301, seed=5).to(DEV)
    outs = []
    monkeypatch.setenv("NOVIC_ATTN_SPLIT", "0")  # small batches otherwise use attention_split_kernel, which sums the keys in another order
    for flag in ("0", "1"):                      # the switch is read when a handle is created
        monkeypatch.setenv("NOVIC_FUSE_ATTN", flag)
        m = default_decoder(dims, sd).to(DEV)
        with torch.inference_mode():
            g = m.generate(embed, True, True, 0.9, 0.2, None, None, False)
            b = m.generate_beam(embed[:70], 3, 1.0, 0.0, None, False, 0.0, None, False)
            small = [m.generate(embed[:n], False, True, 1.0, 0.0, None, None, False) for n in (1, 17, 33)]
        outs.append((g, b, small))
        del m
    (g0, b0, s0), (g1, b1, s1) = outs
    assert torch.equal(g0[0], g1[0]) and torch.equal(g0[1], g1[1]) and torch.equal(g0[2], g1[2]) and torch.equal(g0[5], g1[5])
    assert torch.equal(b0[0], b1[0]) and torch.equal(b0[1], b1[1]) and torch.equal(b0[2], b1[2])
    for a, c in zip(s0, s1):
        assert torch.equal(a[0], c[0]) and torch.equal(a[1], c[1]) and torch.equal(a[5], c[5])


def test_beam_logits_on_wide_tiles_are_bit_identical(monkeypatch):
    """Beam search with up to three beams ranks four candidates per 64-column vocabulary slice in the logits epilogue (EpiLogits<4>).  From
    ~700 candidate rows per step on, that GEMM runs on 128 x 256 tiles with 16 epilogue warps like the greedy one (NOVIC_LOGITS_BN=128
    restores the 128 x 128 tiles with 8 warps): the same products in the same order and the same per-slice records -> bit-identical
    tokens, padding and scores.  300 embeddings x 3 beams = 900 rows per step (a ragged eighth row block); also ten beams (EpiLogits<12>)
    and guided decoding (the masked epilogue), with and without guide_renorm."""
    dims = synth.DecoderDims()
    sd = weight_case("lively")
    embed = synth.synth_embeddings(300, seed=11).to(DEV)
    guide = synth.synth_guide_targets(400, dims, seed=33, first_pool=200).to(DEV)
    outs = []
    for bn in ("128", "256"):                    # the switch is read when a handle is created
        monkeypatch.setenv("NOVIC_LOGITS_BN", bn)
        m = default_decoder(dims, sd).to(DEV)
        with torch.inference_mode():
            runs = [m.generate_beam(embed, 3, 1.0, 0.0, None, False, 0.0, None, False),
                    m.generate_beam(embed, 3, 1.3, 0.6, None, False, 0.0, None, False),
                    m.generate_beam(embed[:90], 10, 1.0, 0.0, None, False, 0.0, None, False),          # 12 candidates per slice, 900 rows
                    m.generate_beam(embed, 3, 1.0, 0.0, None, False, 0.0, guide, False),             # guided: masked epilogue
                    m.generate_beam(embed[:90], 10, 1.0, 0.0, None, False, 0.0, guide, True)]        # the reference's default configuration
        outs.append(runs)
        del m
    for x, y in zip(outs[0], outs[1]):
        assert torch.equal(x[0], y[0]) and torch.equal(x[1], y[1]) and torch.equal(x[2], y[2])


def test_qkv_projection_variants_are_bit_identical(monkeypatch):
    """The QKV projection of a large decode step (>= 24 row blocks) has five implementations with the same operands and accumulation order:
    the generic persistent kernel (NOVIC_QKV_WS=0), the weight-stationary kernel (default), its cluster-multicast variants
    (NOVIC_QKV_MC=2 / 4: activation stages requested once per cluster) and the weight-stationary CTA-pair kernel (NOVIC_QKV_WS=2,
    tcgen05.mma.cta_group::2).  3100 embeddings = 24 full row blocks + a ragged one of 28 rows (the pair kernel's last pair has an empty
    second CTA).  Ids, padding and scores must be bit-identical."""
    dims = synth.DecoderDims()
    sd = weight_case("lively")
    embed = synth.synth_embeddings(3100, seed=8).to(DEV)
    outs = []
    for ws, mc, st in (("0", "0", "2"), ("1", "0", "2"), ("1", "2", "2"), ("1", "4", "2"), ("2", "0", "2"), ("1", "0", "3")):     # the switches are read when a handle is created
        monkeypatch.setenv("NOVIC_QKV_WS", ws)
        monkeypatch.setenv("NOVIC_QKV_MC", mc)
        monkeypatch.setenv("NOVIC_QKV_WS_STAGES", st)     # 3: a third activation stage instead of the staging tile, q / K / V stored with 256-bit stores
        m = default_decoder(dims, sd).to(DEV)
        with torch.inference_mode():
            g = m.generate(embed, False, True, 0.9, 0.2, None, None, False)
        outs.append((g[0].clone(), g[1].clone(), g[5].clone()))
        del m
    for o in outs[1:]:
        assert torch.equal(outs[0][0], o[0]) and torch.equal(outs[0][1], o[1]) and torch.equal(outs[0][2], o[2])
