"""novic_b200.serve.GenerationPipeline: pipelined host -> device -> host inference returns exactly what batch-by-batch
generate / generate_beam calls return (infer.py:556-611 shapes), in order."""
import pytest
import torch

from novic_b200 import default_decoder, synth
from novic_b200.serve import GenerationPipeline

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def model():
    dims = synth.DecoderDims()
    return default_decoder(dims, synth.synth_state_dict(dims, seed=2, token_scale=0.25, jitter_norms=True)).to(DEV)


def host_batches():
    return [synth.synth_embeddings(n, seed=50 + i).pin_memory() for i, n in enumerate((96, 33, 128, 1))]


def test_greedy_pipeline_equals_sequential_calls(model):
    batches = host_batches()
    got = list(GenerationPipeline(model, "greedy", temperature=0.8, length_alpha=0.3).run(batches))
    assert len(got) == len(batches)
    with torch.inference_mode():
        for hb, (tok, pad, score) in zip(batches, got):
            t, p, _, _, _, s = model.generate(hb.to(DEV), False, True, 0.8, 0.3, None, None, False)
            assert not tok.is_cuda and tok.shape == (hb.shape[0], 1, t.shape[1])
            assert torch.equal(tok[:, 0], t.cpu()) and torch.equal(pad[:, 0], p.cpu()) and torch.equal(score[:, 0], s.cpu())


def test_beam_pipeline_with_guide_equals_sequential_calls(model):
    batches = host_batches()[:3]
    gt = synth.synth_guide_targets(300, synth.DecoderDims(), seed=21, first_pool=24).to(DEV)
    got = list(GenerationPipeline(model, "beam", topk=3, guide_targets=gt).run(batches))
    with torch.inference_mode():
        for hb, (tok, pad, score) in zip(batches, got):
            t, p, s = model.generate_beam(hb.to(DEV), 3, 1.0, 0.0, None, False, 0.0, gt, False)
            assert torch.equal(tok, t.cpu()) and torch.equal(pad, p.cpu()) and torch.equal(score, s.cpu())


def test_pipeline_edge_cases(model):
    assert list(GenerationPipeline(model).run([])) == []
    assert list(GenerationPipeline(model, emit=False).run(host_batches()[:2])) == [None, None]
    with pytest.raises(ValueError):
        GenerationPipeline(model, "sample")
