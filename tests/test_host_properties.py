"""Property tests (hypothesis, CPU) of the host-side data structures: the guide / vocabulary trie, the per-edge vocabulary prior, the
id-space membership sets of the statistics and the cache file round trip - each against a brute-force statement of the reference
semantics on small random inputs."""
import os

import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from novic_b200 import cache, guide, stats

V, CMAX = 12, 5
settings.register_profile("novic", max_examples=60, deadline=None)
settings.load_profile("novic")


@st.composite
def noun_sets(draw, min_rows=1, max_rows=12):
    """Unique rows of 1..CMAX-1 ids in [1, V) followed by the end token 0 and zero padding (what infer.py:687-710 produces)."""
    n = draw(st.integers(min_rows, max_rows))
    rows = draw(st.lists(st.lists(st.integers(1, V - 1), min_size=1, max_size=CMAX - 1), min_size=n, max_size=n, unique_by=tuple))
    out = torch.zeros((len(rows), CMAX), dtype=torch.int64)
    for i, r in enumerate(rows):
        out[i, :len(r)] = torch.tensor(r)
    return out


def continuations(targets: torch.Tensor, prefix: list) -> dict:
    """token -> number of targets matching `prefix` that continue with it (brute force, embedding_decoder.py:873-878 / :915-917)."""
    c = len(prefix)
    match = (targets[:, :c] == torch.tensor(prefix, dtype=torch.int64).view(1, -1)).all(dim=1) if c else torch.ones(targets.shape[0], dtype=torch.bool)
    ids, counts = np.unique(targets[match, c].numpy(), return_counts=True)
    return dict(zip(ids.tolist(), counts.tolist()))


@given(noun_sets())
def test_trie_nodes_are_exactly_the_matching_target_sets(gt):
    G = CMAX - 1
    trie = guide.build_trie(gt, G, V)
    off, tok, node = trie.child_off.numpy(), trie.child_tok.numpy(), trie.child_node.numpy()
    assert off[0] == 0 and off[-1] == trie.num_edges and (np.diff(off) >= 0).all()
    for w in range(gt.shape[0]):
        n, prefix = 0, []
        for c in range(G):
            want = continuations(gt, prefix)
            kids = tok[off[n]:off[n + 1]]
            assert kids.tolist() == sorted(want)                               # children = allowed next ids, ascending
            assert trie.child_count.numpy()[off[n]:off[n + 1]].tolist() == [want[k] for k in kids.tolist()]
            assert int(trie.node_count[n]) == sum(want.values())
            t = int(gt[w, c])
            n = int(node[off[n] + int(np.searchsorted(kids, t))])
            prefix.append(t)


@given(noun_sets(), noun_sets(), st.booleans(), st.floats(0.1, 2.0))
def test_prior_bias_of_a_guide_against_another_vocabulary(gt, vt, per_token, scaler):
    G = CMAX - 1
    trie, bias = guide.prior_bias(guide.build_trie(gt, G, V), vt, False, per_token, scaler, G, V)
    off, tok, node = trie.child_off.numpy(), trie.child_tok.numpy(), trie.child_node.numpy()
    for w in range(gt.shape[0]):
        n, prefix = 0, []
        for c in range(G):
            want = continuations(vt, prefix)
            for e in range(off[n], off[n + 1]):
                t = int(tok[e])
                if t in want:
                    p = 1.0 / len(want) if per_token else want[t] / sum(want.values())
                    assert abs(bias[e].item() - (-scaler * np.log(p))) < 1e-4
                else:
                    assert bias[e].item() == float("-inf")
            t = int(gt[w, c])
            n = int(node[off[n] + int(np.searchsorted(tok[off[n]:off[n + 1]], t))])
            prefix.append(t)


@given(noun_sets(), noun_sets(max_rows=20))
def test_row_set_membership_equals_exact_row_match(members, queries):
    G = CMAX - 1
    rs = stats._RowSet(members, G, V, "cpu")
    got = rs.contains(queries[:, :G])
    want = torch.tensor([any(torch.equal(q[:G], m[:G]) for m in members) for q in queries])
    assert torch.equal(got, want)
    assert rs.contains(members[:, :G]).all()
    assert rs.contains(members[:, :2]).tolist() == [bool((members[:, :G] == torch.cat((m[:2], torch.zeros(G - 2, dtype=torch.int64)))).all(dim=1).any()) for m in members]


@given(st.integers(1, 9), st.integers(1, 6), st.integers(1, 4), st.integers(2, 6), st.integers(1, 3), st.booleans(), st.sampled_from([torch.int16, torch.int32, torch.int64]))
def test_cache_file_round_trip(N, F, R1, C, M, with_weights, id_dtype):
    g = torch.Generator().manual_seed(N * 131 + F * 17 + R1 * 7 + C * 3 + M)
    emb = torch.nn.functional.normalize(torch.randn(N, F, generator=g), dim=-1)
    tok = torch.randint(1, 50, (R1, C), dtype=torch.int64, generator=g)
    lens = torch.randint(0, C - 1, (R1,), generator=g)
    mask = torch.arange(C).unsqueeze(0) > lens.unsqueeze(1)
    tok[mask] = 0
    et = torch.randint(1, R1 + 1, (N, M), generator=g).to(id_dtype)
    if M > 1:
        et[::2, M - 1] = 0
    w = None
    if with_weights:
        w = torch.rand(N, M, generator=g) + 0.1
        w[et == 0] = 0
        w = w / w.sum(dim=1, keepdim=True)
    path = os.path.join(os.environ.get("TMPDIR", "/tmp"), f"novic_cache_prop_{os.getpid()}.bin")
    try:
        h = cache.write_cache(path, emb, [f"n{i}" for i in range(R1)], tok, mask, et, w)
        assert h.embed_targets_dtype_id == cache.INT_DTYPES.index(id_dtype) and h.default_weights == (w is None)
        with cache.EmbeddingCacheReader(path, device="cpu") as r:
            e, ids, t, m, wt = r.get_samples(0, N)
            assert torch.equal(e, emb) and torch.equal(ids, et)
            full_tok = torch.cat((torch.zeros(1, C, dtype=torch.int64), tok))
            full_mask = torch.cat((torch.ones(1, C, dtype=torch.bool), mask))
            assert torch.equal(t, full_tok[et.long()]) and torch.equal(m, full_mask[et.long()])
            if w is not None:
                assert torch.allclose(wt, w)
            assert sum(b[0].shape[0] for b in r.batches(min(N, 4))) == N
            assert sum(b[0].shape[0] for b in r.batches(min(N, 4), training=True, epoch_index_offset=1)) == (N // min(N, 4)) * min(N, 4)
    finally:
        if os.path.exists(path):
            os.remove(path)


def test_grad_bucket_recognises_only_its_own_views():
    """dist.GradBucket.covers: the in-place all-reduce may only run when the gradients are exactly the bucket's views."""
    from novic_b200.dist import GradBucket
    flat = torch.zeros(12 + 5 + GradBucket.SPARE)
    views = [flat[:12].view(3, 4), flat[12:17]]
    b = GradBucket(flat, views)
    assert b.total == 17 and b.covers(views) and b.covers(list(reversed(views)))          # order does not matter, memory does
    assert b.covers([v.detach() for v in views])                                            # autograd hands over detached aliases
    assert not b.covers(views[:1])                                                          # a gradient is missing
    assert not b.covers([views[0], views[0]])                                               # one view twice
    assert not b.covers([views[0], views[1].clone()])                                       # a copy lives elsewhere
    assert not b.covers([views[0], flat[12:16]])                                            # right address, wrong extent
    assert not b.covers([views[0].double(), views[1]])                                      # cast gradients
    assert not b.covers([flat[:12].view(4, 3).t(), views[1]])                               # same memory, not contiguous
