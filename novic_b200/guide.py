"""Host side of guided decoding: the reference's W x Cmax `guide_targets` tensor (infer.py:687-710) as a token trie.

The reference tracks, for every candidate sequence, which of the W guide targets still match the generated prefix
(a B x H x W boolean mask, embedding_decoder.py:873-878) and scatters their next token ids into a B x H x (V+1) score
tensor at every step (:915-917).  The set of guide targets that match a prefix is exactly a trie node, and the token ids
allowed next are the node's child edges, so the CUDA path carries one int32 node id per sequence instead.

Layout handed to the library (include/novic_b200.h, struct NovicGuide): CSR over nodes, node 0 = root,
`child_tok[child_off[n] : child_off[n + 1]]` ascending, `child_node` the node each edge leads to.  Nodes are numbered
depth by depth in lexicographic order of their prefixes.
"""
from __future__ import annotations

import ctypes as C
import dataclasses

import numpy as np
import torch

from . import _abi


@dataclasses.dataclass
class GuideTrie:
    child_off: torch.Tensor   # int32 [num_nodes + 1]
    child_tok: torch.Tensor   # int32 [num_edges]
    child_node: torch.Tensor  # int32 [num_edges]
    num_nodes: int
    num_edges: int
    depth: int
    child_count: torch.Tensor = None   # int64 [num_edges] guide targets below each child (host only; vocabulary priors)
    node_count: torch.Tensor = None    # int64 [num_nodes] guide targets below each node (host only)

    def to(self, device) -> "GuideTrie":
        return GuideTrie(self.child_off.to(device), self.child_tok.to(device), self.child_node.to(device), self.num_nodes, self.num_edges,
                         self.depth, self.child_count, self.node_count)

    def as_struct(self, renorm: bool, bias: torch.Tensor = None) -> _abi.NovicGuide:
        if bias is not None:
            assert bias.dtype == torch.float32 and bias.numel() == self.num_edges and bias.device == self.child_off.device
        return _abi.NovicGuide(self.child_off.data_ptr(), self.child_tok.data_ptr(), self.child_node.data_ptr(), self.num_nodes,
                               self.num_edges, 1 if renorm else 0, None if bias is None else bias.data_ptr())


def build_trie(guide_targets: torch.Tensor, gen_len: int, vocab_size: int) -> GuideTrie:
    """guide_targets: W x Cmax token ids (end token / padding = 0).  Only the first gen_len = Cmax - 1 positions can be
    generated (embedding_decoder.py:782), so the trie has gen_len levels below the root."""
    gt = guide_targets.detach().cpu().numpy().astype(np.int64)
    if gt.ndim != 2 or gt.shape[0] < 1:
        raise ValueError("guide_targets must be a non-empty W x Cmax tensor")
    if gt.min() < 0 or gt.max() >= vocab_size:
        raise ValueError("guide_targets contains token ids outside [0, vocab_size)")
    G = min(gen_len, gt.shape[1])
    gt = gt[:, :G]
    W = gt.shape[0]
    order = np.lexsort(gt.T[::-1])                 # rows in lexicographic order (first column is the primary key)
    srt = gt[order]
    # new[i, d]: row i starts a new prefix of length d (d = 0: only the first row, the root)
    new = np.zeros((W, G + 1), dtype=bool)
    new[0, :] = True
    if W > 1:
        differs = np.maximum.accumulate(srt[1:] != srt[:-1], axis=1)   # differs[i, d]: rows i, i+1 differ somewhere in columns <= d
        new[1:, 1:] = differs
    counts = new.sum(axis=0)                        # nodes per depth
    base = np.concatenate(([0], np.cumsum(counts)))
    num_nodes = int(base[-1])
    nid = np.cumsum(new, axis=0) - 1 + base[:-1][None, :]              # node id of row i's prefix of length d
    parents, toks, childs = [], [], []
    for d in range(G):
        rows = np.nonzero(new[:, d + 1])[0]         # one row per node of depth d + 1, already ordered by (parent, token)
        parents.append(nid[rows, d]); toks.append(srt[rows, d]); childs.append(nid[rows, d + 1])
    parent = np.concatenate(parents) if parents else np.zeros(0, dtype=np.int64)
    tok = np.concatenate(toks) if toks else np.zeros(0, dtype=np.int64)
    child = np.concatenate(childs) if childs else np.zeros(0, dtype=np.int64)
    child_off = np.zeros(num_nodes + 1, dtype=np.int64)
    np.add.at(child_off, parent + 1, 1)
    child_off = np.cumsum(child_off)
    # rows below each node: consecutive in the sorted order, so the count is the distance to the next node of the same depth
    node_count = np.zeros(num_nodes, dtype=np.int64)
    for d in range(G + 1):
        starts = np.nonzero(new[:, d])[0]
        node_count[base[d]:base[d + 1]] = np.diff(np.concatenate((starts, [W])))
    return GuideTrie(torch.from_numpy(child_off.astype(np.int32)), torch.from_numpy(tok.astype(np.int32)),
                     torch.from_numpy(child.astype(np.int32)), num_nodes, int(tok.shape[0]), G,
                     torch.from_numpy(node_count[child]), torch.from_numpy(node_count))


def target_paddings(guide_targets: torch.Tensor) -> torch.Tensor:
    """W x Cmax bool: position c is padding iff an earlier position holds the end token 0 (embedding_decoder.py:991-994)."""
    pads = torch.zeros_like(guide_targets, dtype=torch.bool)
    pads[:, 1:] = (guide_targets[:, :-1] == 0).cummax(dim=1).values
    return pads


def vocab_prior_scores(targets: torch.Tensor, paddings: torch.Tensor, vocab_targets: torch.Tensor, per_token: bool,
                       vocab_size: int) -> torch.Tensor:
    """Sum over the unpadded positions of log p_vocab(target token | target prefix) for each of the W targets
    (embedding_decoder.py:1020-1044, before the vocab_scaler factor): p_vocab is uniform over the distinct continuations of
    the vocabulary nouns sharing the prefix (per_token) or proportional to how many of them take each continuation.
    A target that leaves the vocabulary trie gets +inf, like the reference's nan_to_num(log 0)."""
    W, C = targets.shape
    trie = build_trie(vocab_targets, C, vocab_size)
    off = trie.child_off.numpy().astype(np.int64)
    keys = np.repeat(np.arange(trie.num_nodes, dtype=np.int64), np.diff(off)) * vocab_size + trie.child_tok.numpy().astype(np.int64)
    child = trie.child_node.numpy().astype(np.int64)
    ccount, ncount = trie.child_count.numpy().astype(np.float64), trie.node_count.numpy().astype(np.float64)
    nchild = np.diff(off).astype(np.float64)
    tg, pd = targets.cpu().numpy().astype(np.int64), paddings.cpu().numpy()
    node = np.zeros(W, dtype=np.int64)
    alive = np.ones(W, dtype=bool)
    total = np.zeros(W, dtype=np.float64)
    for c in range(min(C, trie.depth)):
        q = node * vocab_size + tg[:, c]
        idx = np.minimum(np.searchsorted(keys, q), len(keys) - 1) if len(keys) else np.zeros(W, dtype=np.int64)
        found = alive & (len(keys) > 0) & (keys[idx] == q)
        p = np.where(found, (1.0 / np.maximum(nchild[node], 1.0)) if per_token else ccount[idx] / np.maximum(ncount[node], 1.0), 0.0)
        with np.errstate(divide="ignore"):
            lp = np.where(p > 0, np.log(np.maximum(p, 1e-300)), np.inf)
        total += np.where(pd[:, c], 0.0, lp)
        alive = found
        node = np.where(found, child[idx], 0)
    return torch.from_numpy(total.astype(np.float32))


def _edge_logp(trie: GuideTrie, per_token: bool) -> np.ndarray:
    """log p(child | node) for every edge of a trie: uniform over the node's distinct continuations (per_token), else the
    fraction of the node's targets that take the edge (embedding_decoder.py:926-933)."""
    off = trie.child_off.numpy().astype(np.int64)
    parent = np.repeat(np.arange(trie.num_nodes, dtype=np.int64), np.diff(off))
    if per_token:
        p = 1.0 / np.diff(off)[parent].astype(np.float64)
    else:
        p = trie.child_count.numpy().astype(np.float64) / trie.node_count.numpy().astype(np.float64)[parent]
    return np.log(p)


def prior_bias(guide_trie, vocab_targets: torch.Tensor, vocab_is_guide: bool, per_token: bool, scaler: float, gen_len: int,
               vocab_size: int):
    """Vocabulary prior of the beam search (embedding_decoder.py:924-936) as one additive term per trie edge:
    -scaler * log p_vocab(token | prefix), or -inf for a continuation no vocabulary noun takes.  Returns (trie, bias[num_edges]):
    the guide trie when decoding is guided (its edges looked up in the vocabulary trie along the same prefixes), else the
    vocabulary trie itself - without a guide the prior alone restricts decoding to the vocabulary nouns."""
    if guide_trie is None or vocab_is_guide:
        trie = guide_trie if guide_trie is not None else build_trie(vocab_targets, gen_len, vocab_size)
        return trie, torch.from_numpy((-scaler * _edge_logp(trie, per_token)).astype(np.float32))
    vt = build_trie(vocab_targets, gen_len, vocab_size)
    v_off = vt.child_off.numpy().astype(np.int64)
    v_keys = np.repeat(np.arange(vt.num_nodes, dtype=np.int64), np.diff(v_off)) * vocab_size + vt.child_tok.numpy().astype(np.int64)
    v_child = vt.child_node.numpy().astype(np.int64)
    v_logp = _edge_logp(vt, per_token)
    g_off = guide_trie.child_off.numpy().astype(np.int64)
    g_parent = np.repeat(np.arange(guide_trie.num_nodes, dtype=np.int64), np.diff(g_off))
    g_tok = guide_trie.child_tok.numpy().astype(np.int64)
    g_child = guide_trie.child_node.numpy().astype(np.int64)
    vmap = np.full(guide_trie.num_nodes, -1, dtype=np.int64)    # guide node -> vocabulary node with the same prefix
    vmap[0] = 0
    logp = np.full(guide_trie.num_edges, -np.inf)
    # walk both tries level by level: the vocabulary node of a guide node's prefix, and the prior of every guide edge
    depth = np.zeros(guide_trie.num_nodes, dtype=np.int64)      # node depth: parents precede children, one relaxation per level
    for _ in range(guide_trie.depth):
        depth[g_child] = depth[g_parent] + 1
    for d in range(guide_trie.depth):
        es = np.nonzero(depth[g_parent] == d)[0]
        if es.size == 0:
            continue
        vp = vmap[g_parent[es]]
        q = np.where(vp >= 0, vp, 0) * vocab_size + g_tok[es]
        idx = np.minimum(np.searchsorted(v_keys, q), len(v_keys) - 1) if len(v_keys) else np.zeros(es.size, dtype=np.int64)
        found = (vp >= 0) & (len(v_keys) > 0) & (v_keys[idx] == q)
        vmap[g_child[es]] = np.where(found, v_child[idx], -1)
        logp[es] = np.where(found, v_logp[idx], -np.inf)
    with np.errstate(invalid="ignore"):
        bias = np.where(np.isfinite(logp), -scaler * logp, -np.inf)
    return guide_trie, torch.from_numpy(bias.astype(np.float32))


def tensor_version(t: torch.Tensor) -> int:
    try:
        return t._version
    except RuntimeError:                # inference tensors do not track versions
        return -1


def content_checksum(t: torch.Tensor) -> int:
    """Position-weighted 64-bit checksum of an integer tensor (wrap-around arithmetic on the tensor's device, one scalar read
    back).  Used where a version counter cannot tell whether a guide tensor changed: inference tensors have none, yet they can be
    written in place inside torch.inference_mode()."""
    flat = t.reshape(-1).to(torch.int64)
    w = torch.arange(1, flat.numel() + 1, device=flat.device, dtype=torch.int64) * 0x9E3779B1
    return int(((flat + 1) * w).sum().item())


class TrieCache:
    """Tries keyed by identity, version and (for tensors without a version counter) content of the guide tensor - infer.py hands the
    same tensor to every batch.  An entry keeps a strong reference to its source tensor: while the entry lives the tensor's
    storage cannot be freed, so a different tensor can never appear at the cached address (address-only keys return a
    stale trie after the allocator reuses a freed guide tensor's block)."""

    def __init__(self, max_entries: int = 8):
        self.entries: list[tuple[tuple, torch.Tensor, GuideTrie]] = []
        self.max_entries = max_entries

    def get(self, guide_targets: torch.Tensor, gen_len: int, vocab_size: int, device, check_content: bool = True) -> GuideTrie:
        version = tensor_version(guide_targets)
        ident = (guide_targets.data_ptr(), tuple(guide_targets.shape), tuple(guide_targets.stride()), version, str(guide_targets.device),
                 str(device), gen_len, vocab_size)
        # a version of -1 says nothing: compare contents instead (skipped only by callers that must not synchronise)
        checksum = content_checksum(guide_targets) if (version < 0 and check_content) else None
        for i, (k, src, trie) in enumerate(self.entries):
            if k[0] == ident and (checksum is None or k[1] is None or k[1] == checksum):
                if i != len(self.entries) - 1:                      # most recently used last
                    self.entries.append(self.entries.pop(i))
                return trie
        trie = build_trie(guide_targets, gen_len, vocab_size).to(device)
        self.entries = [e for e in self.entries if e[0][0] != ident]   # same tensor, new content: drop the stale entry
        self.entries.append(((ident, checksum), guide_targets, trie))
        if len(self.entries) > self.max_entries:
            self.entries.pop(0)
        return trie

    def clear(self) -> None:
        self.entries.clear()


def guide_arg(trie, renorm: bool, bias: torch.Tensor = None):
    """ctypes argument for the `const NovicGuide*` parameter (None = unguided)."""
    return None if trie is None else C.byref(trie.as_struct(renorm, bias))
