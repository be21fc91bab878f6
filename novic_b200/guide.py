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

    def to(self, device) -> "GuideTrie":
        return GuideTrie(self.child_off.to(device), self.child_tok.to(device), self.child_node.to(device), self.num_nodes, self.num_edges, self.depth)

    def as_struct(self, renorm: bool) -> _abi.NovicGuide:
        return _abi.NovicGuide(self.child_off.data_ptr(), self.child_tok.data_ptr(), self.child_node.data_ptr(), self.num_nodes,
                               self.num_edges, 1 if renorm else 0)


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
    return GuideTrie(torch.from_numpy(child_off.astype(np.int32)), torch.from_numpy(tok.astype(np.int32)),
                     torch.from_numpy(child.astype(np.int32)), num_nodes, int(tok.shape[0]), G)


class TrieCache:
    """Tries keyed by the identity and version of the guide tensor (infer.py hands the same tensor to every batch)."""

    def __init__(self, max_entries: int = 4):
        self.entries: list[tuple[tuple, GuideTrie]] = []
        self.max_entries = max_entries

    def get(self, guide_targets: torch.Tensor, gen_len: int, vocab_size: int, device) -> GuideTrie:
        try:
            version = guide_targets._version
        except RuntimeError:            # inference tensors do not track versions; they cannot be modified in place either
            version = -1
        key = (guide_targets.data_ptr(), tuple(guide_targets.shape), version, str(guide_targets.device), str(device), gen_len, vocab_size)
        for k, trie in self.entries:
            if k == key:
                return trie
        trie = build_trie(guide_targets, gen_len, vocab_size).to(device)
        self.entries.append((key, trie))
        if len(self.entries) > self.max_entries:
            self.entries.pop(0)
        return trie


def guide_arg(trie, renorm: bool):
    """ctypes argument for the `const NovicGuide*` parameter (None = unguided)."""
    return None if trie is None else C.byref(trie.as_struct(renorm))
