"""SURVEY.md section 8 row a13 on hardware: the reference's OWN glue code - infer.load_decoder_model (infer.py:713-778),
GenerationConfig.from_name (:374-433), GenerationTask.generate / update / process (:556-644) - runs unmodified on the registered
CUDA class, and returns what direct calls of the class return.

The reference modules come from oracle/_ref (byte-for-byte staged by oracle/build_ref.py; /root/reference does not exist on the
GPU box).  The CLIP tokenizer is not available offline, so `detokenize_target` is an injective stand-in (ids up to the end token).
"""
import types

import pytest
import torch

import novic_b200
from novic_b200 import stats, synth
from oracle import refload
from tests.golden_util import weight_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
DIMS = synth.DecoderDims()
G = DIMS.token_length - 1


def detok(rows: torch.Tensor):
    def one(r):
        out = []
        for t in r.tolist():
            if t == 0:
                break
            out.append(str(t))
        return " ".join(out)
    return [[one(r) for r in b] for b in rows]


@pytest.fixture(scope="module")
def ref():
    if not refload.available():
        pytest.skip("no staged reference (python oracle/build_ref.py in the build container)")
    return refload.import_reference()


@pytest.fixture(scope="module")
def loaded(ref):
    """The decoder exactly as infer.load_decoder_model builds it once novic_b200.register() has been called."""
    original = ref.embedding_decoder.PrefixedIterDecoder
    novic_b200.register(ref.embedding_decoder)
    try:
        model = refload.build_reference_decoder(ref, weight_case("eos"))
    finally:
        ref.embedding_decoder.PrefixedIterDecoder = original
    assert type(model) is novic_b200.PrefixedIterDecoder
    model.embedder.detokenize_target = detok
    return model.to(DEV)


def make_task(ref, model, name, guide_t, vocab_t, classes=None):
    gencfg = ref.infer.GenerationConfig.from_name(name)
    flat = lambda t: {s for row in detok(t.unsqueeze(0)) for s in row}   # noqa: E731
    return ref.infer.GenerationTask(
        gencfg=gencfg, decoder=model, vocab_targets_set=flat(vocab_t), vocab_targets=vocab_t.to(DEV) if gencfg.vocab_prior else None,
        guide_targets_set=flat(guide_t), guide_targets=guide_t.to(DEV) if gencfg.guided else None,
        class_lists=None if classes is None else [[s for row in detok(c.unsqueeze(0)) for s in row] for c in classes])


@pytest.mark.parametrize("name", ["greedy_k1_vnone_gn_t1_a0", "beam_k10_vnone_gp_t1_a0", "beam_k3_vtgt0.5_gr_t0.9_a0.3", "greedy_k1_vnone_gr_t0.8_a0.3",
                                  "all_k5_vnone_gr_t1_a0"])
def test_generation_task_runs_on_the_cuda_class(ref, loaded, name):
    guide_t = synth.synth_guide_targets(400, DIMS, seed=21, first_pool=24)
    vocab_t = torch.cat((guide_t[:250], synth.synth_guide_targets(200, DIMS, seed=25, first_pool=24)))
    g = torch.Generator().manual_seed(5)
    classes = [guide_t[torch.randint(0, 400, (3,), generator=g)] for _ in range(6)]
    embed = synth.synth_embeddings(48, seed=1234).to(DEV)
    class_idx = torch.randint(0, 6, (48,), generator=g).tolist()
    task = make_task(ref, loaded, name, guide_t, vocab_t, classes)
    cfg = task.gencfg
    mine = stats.GenerationStats(cfg.topk, vocab_t, guide_t, G, DIMS.vocab_size, device=DEV, class_targets=classes)
    with torch.inference_mode():
        for lo, hi in ((0, 16), (16, 48)):                      # two batches: GenerationTask accumulates its counters
            task.process(embed[lo:hi], class_indices=class_idx[lo:hi])          # infer.py:512-516 -> :556-611 -> :613-644
            # direct calls of the class with the arguments GenerationTask.generate assembles
            gt = guide_t.to(DEV) if cfg.guided else None
            if cfg.method == "greedy":
                t, p, _, _, _, s = loaded.generate(embed[lo:hi], False, True, cfg.temperature, cfg.length_alpha, None, gt, cfg.guide_renorm)
                t, p, s = t.unsqueeze(1), p.unsqueeze(1), s.unsqueeze(1)
            elif cfg.method == "beam":
                t, p, s = loaded.generate_beam(embed[lo:hi], cfg.topk, cfg.temperature, cfg.length_alpha, vocab_t.to(DEV) if cfg.vocab_prior else None,
                                               cfg.vocab_per_token, cfg.vocab_scaler, gt, cfg.guide_renorm)
            else:
                t, p, s = loaded.generate_all(embed[lo:hi], cfg.topk, cfg.temperature, cfg.length_alpha, None, cfg.vocab_per_token, cfg.vocab_scaler,
                                              gt, cfg.guide_renorm)
            assert task.target.shape == (hi - lo, cfg.topk, t.shape[2])
            assert torch.equal(task.target, t.cpu()) and torch.equal(task.target_padding, p.cpu())        # deterministic: same call, same ids
            assert torch.allclose(torch.tensor(task.target_score), s.cpu(), atol=1e-6)
            # and the reference's string statistics equal the id statistics computed on the device (row f4)
            mine.update(t, p, s, class_indices=class_idx[lo:hi])
            assert torch.equal(mine.valid_guide.cpu(), task.valid_guide) and torch.equal(mine.valid_vocab.cpu(), task.valid_vocab)
            assert torch.equal(mine.correct.cpu(), task.correct) and torch.equal(mine.result.cpu(), task.result)
            assert torch.equal(mine.topk_counts.cpu(), task.topk_counts)
    if cfg.guided:
        assert task.valid_guide.all()                            # every guided prediction spells a guide noun
    assert task.num_samples == 48


def test_reference_decoder_on_the_host_agrees_with_the_cuda_class(ref, loaded):
    """The unmodified reference class on the host cores of this box against the CUDA class on the same inputs - the live version of
    the committed fixtures (ids identical wherever the reference's top-2 margin exceeds 0.12, logits within 0.06)."""
    theirs = refload.build_reference_decoder(ref, weight_case("eos"))
    embed = synth.synth_embeddings(24, seed=77)
    with torch.inference_mode():
        rt, rp, rl, _, _, rs = theirs.generate(embed, True, True, 1.0, 0.0, None, None, False)
        t, p, lg, _, _, s = loaded.generate(embed.to(DEV), True, True, 1.0, 0.0, None, None, False)
    t, p, lg, s = t.cpu(), p.cpu(), lg.cpu(), s.cpu()
    top2 = rl.topk(2, dim=-1).values
    margin = top2[..., 0] - top2[..., 1]
    first = rl[:, 0, 1:].topk(2, dim=-1).values
    margin[:, 0] = first[:, 0] - first[:, 1]
    n = min(t.shape[1], rt.shape[1])
    diff = t[:, :n] != rt[:, :n]
    first = torch.where(diff.any(dim=1), diff.float().argmax(dim=1), torch.full((24,), n))
    clean = first >= n
    for b in (~clean).nonzero().flatten().tolist():      # a row may leave the reference's path only where the reference was undecided
        assert margin[b, first[b]] <= 0.12, f"row {b} diverged at step {first[b]} with margin {margin[b, first[b]]:.3f}"
    assert clean.float().mean() >= 0.9
    assert torch.equal(p[clean, :n], rp[clean, :n])
    live = (~rp[:, :n]) & clean.unsqueeze(1)
    assert (lg[:, :n] - rl[:, :n])[live].abs().max() <= 0.06
    n_tok = (~rp).sum(dim=1).float()
    assert ((s - rs).abs()[clean] <= 2 * 0.06 * n_tok[clean].clamp(min=1) + 1e-3).all()
