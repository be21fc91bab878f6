"""world_size-2 gloo test of the multi-GPU inference plumbing (contiguous shards + final gather), on CPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from novic_b200.dist import generate_sharded, shard_bounds


def test_shard_bounds_partition():
    for n in (0, 1, 7, 8, 4096, 65537):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


class _FakeDecoder:
    """Stands in for the CUDA decoder on CPU: deterministic outputs derived from the embedding rows, and a T that
    depends on the shard (early exit is per shard)."""

    def generate(self, embed, collect_logits, calc_loss, temperature, length_alpha, sample_weight, guide_targets, guide_renorm):
        key = (embed[:, 0] * 1000).round().long()
        T = 3 if int(key.min()) < 5 else 5
        tok = (key.unsqueeze(1) + torch.arange(T)).clamp(min=1)
        pad = torch.zeros_like(tok, dtype=torch.bool)
        return tok, pad, None, None, None, key.float()

    def generate_beam(self, embed, topk, temperature, length_alpha, vocab_targets, vocab_per_token, vocab_scaler, guide_targets, guide_renorm):
        tok, pad, _, _, _, score = self.generate(embed, False, True, temperature, length_alpha, None, None, False)
        return (tok.unsqueeze(1).repeat(1, topk, 1), pad.unsqueeze(1).repeat(1, topk, 1), score.unsqueeze(1) - torch.arange(topk))


def _worker(rank, world, port, n, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        embed = torch.arange(n, dtype=torch.float32).unsqueeze(1).repeat(1, 4) / 1000
        tok, pad, score = generate_sharded(_FakeDecoder(), embed, "greedy")
        btok, bpad, bscore = generate_sharded(_FakeDecoder(), embed, "beam", topk=3)
        if rank == 0:
            results.put((tok, pad, score, btok, bpad, bscore))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", (9, 16))
def test_two_rank_gather_matches_single_process(n):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    results = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, results)) for r in range(2)]
    for p in procs:
        p.start()
    tok, pad, score, btok, bpad, bscore = results.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tok.shape == (n, 1, 5) and btok.shape == (n, 3, 5) and score.shape == (n, 1) and bscore.shape == (n, 3)
    key = torch.arange(n)
    assert torch.equal(score[:, 0], key.float())                      # rank order preserved
    lo, hi = shard_bounds(n, 2, 0)
    # rank 0's shard stopped at T = 3: its extra columns are (id 0, padding True), like samples that finished early
    assert (tok[lo:hi, 0, 3:] == 0).all() and pad[lo:hi, 0, 3:].all() and not pad[lo:hi, 0, :3].any()
    assert not pad[hi:].any() and torch.equal(tok[hi:, 0, 0], key[hi:].clamp(min=1))
    assert torch.equal(btok[:, 1], btok[:, 0]) and torch.equal(bscore[:, 2], key.float() - 2)


def _grad_worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from novic_b200.dist import allreduce_gradients
        torch.manual_seed(0)
        params = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(2, 2))]
        params[0].grad = torch.full((5, 3), float(rank + 1))
        params[1].grad = torch.arange(7, dtype=torch.float32) * (rank + 1)
        params[2].grad = None                                   # parameters without a gradient are skipped consistently
        extra = allreduce_gradients(params, torch.tensor([10.0 * (rank + 1), 1.0]))
        # the same sums when the gradients are views of one flat bucket (the training step's layout): the collective runs in place
        from novic_b200.dist import GradBucket
        flat = torch.zeros(15 + 7 + GradBucket.SPARE)
        views = [flat[:15].view(5, 3), flat[15:22]]
        views[0].fill_(float(rank + 1)); views[1].copy_(torch.arange(7, dtype=torch.float32) * (rank + 1))
        bparams = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(2, 2))]
        bparams[0].grad, bparams[1].grad = views[0], views[1]
        bucket = GradBucket(flat, views)
        assert bucket.covers([p.grad for p in bparams if p.grad is not None])
        bextra = allreduce_gradients(bparams, torch.tensor([10.0 * (rank + 1), 1.0]), bucket=bucket)
        in_place = bparams[0].grad.data_ptr() == flat.data_ptr() and torch.equal(flat[:15].view(5, 3), bparams[0].grad)
        # a gradient that is not the bucket's view (e.g. accumulated twice by autograd) must fall back to the copying path
        bparams[1].grad = bparams[1].grad.clone()
        assert not bucket.covers([p.grad for p in bparams if p.grad is not None])
        allreduce_gradients(bparams, None, bucket=bucket)
        if rank == 0:
            results.put((params[0].grad.clone(), params[1].grad.clone(), extra, bparams[0].grad.clone(), bparams[1].grad.clone(), bextra, in_place))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_allreduce():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    results = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, results)) for r in range(2)]
    for p in procs:
        p.start()
    g0, g1, extra, b0, b1, bextra, in_place = results.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert torch.equal(g0, torch.full((5, 3), 3.0))
    assert torch.equal(g1, torch.arange(7, dtype=torch.float32) * 3)
    assert torch.equal(extra, torch.tensor([30.0, 2.0]))
    assert in_place and torch.equal(bextra, torch.tensor([30.0, 2.0]))
    assert torch.equal(b0, torch.full((5, 3), 6.0))                          # summed twice: bucket path (1 + 2), then the fallback path (3 + 3)
    assert torch.equal(b1, torch.arange(7, dtype=torch.float32) * 3 * 2)
