"""Validity statistics of generated nouns on the device (second half): what GenerationTask.update
(infer.py:613-644) computes after every batch, restated on token ids.

The reference reads the B x K x C predictions back, detokenises them to strings with the CLIP tokenizer (embedders.py:387-406) and tests
every string for membership in Python sets (vocabulary nouns, guide nouns, the ground-truth class's synonyms).  The tokenizer is an
un-vendored third-party dependency, and the predictions are already token rows of the same `compact` id space the target tensors use -
so membership is tested here on the ids: a prediction is "in" a set iff its zero-padded token row equals one of the set's rows, decided
by walking the set's token trie (novic_b200/guide.py) with torch.searchsorted on the device, all B x K rows at once.  This equals the
string test whenever tokenisation is injective on the nouns involved (two different id rows never spell the same noun), which holds
for the reference's own target tensors (they are built by tokenising a duplicate-free noun list, infer.py:687-710).

update() then mirrors the reference's bookkeeping exactly: result codes (0 correct, 1 valid guide, 2 valid vocabulary, 3 invalid), the
cumulative top-k counts and the five top-k ratio vectors."""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import guide


class _RowSet:
    """A set of token rows (W x Cmax, end token / padding = 0) as a device-resident trie; contains(rows) -> bool per row."""

    def __init__(self, targets: torch.Tensor, gen_len: int, vocab_size: int, device):
        trie = guide.build_trie(targets, gen_len, vocab_size)
        off = trie.child_off.to(torch.int64)
        parent = torch.repeat_interleave(torch.arange(trie.num_nodes, dtype=torch.int64), off[1:] - off[:-1])
        self.keys = (parent * vocab_size + trie.child_tok.to(torch.int64)).to(device)     # sorted: edges are ordered by (node, token)
        self.child = trie.child_node.to(torch.int64).to(device)
        self.vocab_size, self.depth = vocab_size, trie.depth

    def contains(self, rows: torch.Tensor) -> torch.Tensor:
        """rows: N x T int64 on the device (T <= depth; shorter rows are zero-extended).  True where the row spells a member."""
        N, T = rows.shape
        node = torch.zeros(N, dtype=torch.int64, device=rows.device)
        alive = torch.ones(N, dtype=torch.bool, device=rows.device)
        last = self.keys.numel() - 1
        for c in range(self.depth):
            tok = rows[:, c] if c < T else torch.zeros(N, dtype=torch.int64, device=rows.device)
            q = node * self.vocab_size + tok
            idx = torch.searchsorted(self.keys, q).clamp_(max=last)
            found = alive & (self.keys[idx] == q)
            node = torch.where(found, self.child[idx], node)
            alive = found
        return alive


class GenerationStats:
    """Running top-k statistics over batches of predictions, field for field like GenerationTask (infer.py:445-468, :613-644).

    vocab_targets / guide_targets: Z x Cmax / W x Cmax token tensors (what the reference holds besides its string sets);
    class_targets: optional sequence over classes of n_c x Cmax tensors - the tokenised synonyms that count as correct for class c."""

    def __init__(self, topk: int, vocab_targets: torch.Tensor, guide_targets: torch.Tensor, gen_len: int, vocab_size: int, device="cuda",
                 class_targets: Optional[Sequence[torch.Tensor]] = None):
        self.topk_k = int(topk)
        self.device = torch.device(device)
        self.vocab = _RowSet(vocab_targets, gen_len, vocab_size, self.device)
        self.guide = _RowSet(guide_targets, gen_len, vocab_size, self.device)
        self.classes = None
        if class_targets is not None:
            # one trie over all classes' synonyms with the class index appended as a last pseudo-token: membership of (row, class)
            C = max(t.shape[1] for t in class_targets)
            rows = []
            for ci, t in enumerate(class_targets):
                r = torch.zeros((t.shape[0], gen_len + 1), dtype=torch.int64)
                r[:, :min(gen_len, t.shape[1])] = t[:, :gen_len]
                r[:, gen_len] = ci + 1
                rows.append(r)
            self._class_width = gen_len + 1
            self.classes = _RowSet(torch.cat(rows), gen_len + 1, max(vocab_size, len(class_targets) + 2), self.device)
        self.gen_len = gen_len
        self.clear()

    def clear(self):
        self.num_samples = 0
        self.topk_counts = torch.zeros((self.topk_k, 4), dtype=torch.int64, device=self.device)
        self.valid_vocab = self.valid_guide = self.correct = self.invalid = self.result = None
        self.topk_invalid = self.topk_valid = self.topk_vocab = self.topk_guide = self.topk = None

    def update(self, target: torch.Tensor, target_padding: torch.Tensor, target_score=None, *, class_indices=None):
        """target / target_padding: B x K x T on the device (as generate_beam / generate_all return them).  Mirrors infer.py:623-644."""
        B, K, T = target.shape
        assert K == self.topk_k
        rows = target.masked_fill(target_padding, 0).reshape(B * K, T).to(self.device)
        self.num_samples += B
        self.valid_vocab = self.vocab.contains(rows).view(B, K)
        self.valid_guide = self.guide.contains(rows).view(B, K)
        if class_indices is not None and self.classes is not None:
            ci = torch.as_tensor(class_indices, dtype=torch.int64, device=self.device).view(B, 1).expand(B, K).reshape(B * K, 1) + 1
            ext = torch.zeros((B * K, self._class_width), dtype=torch.int64, device=self.device)
            ext[:, :min(T, self.gen_len)] = rows[:, :self.gen_len]
            ext[:, self.gen_len:] = ci
            self.correct = self.classes.contains(ext).view(B, K)
        else:
            self.correct = torch.zeros((B, K), dtype=torch.bool, device=self.device)
        self.invalid = ~(self.valid_vocab | self.valid_guide | self.correct)
        stacked = torch.stack((self.correct, self.valid_guide, self.valid_vocab, torch.ones_like(self.invalid)), dim=2).cummax(dim=2)[0]
        self.result = torch.max(stacked, dim=2)[1]                                     # infer.py:634
        stacked[:, :, -1] = self.invalid
        self.topk_counts.add_(stacked.cummax(dim=1)[0].sum(dim=0))
        counts = self.topk_counts.to(torch.float32)
        self.topk_valid = (self.num_samples - counts[:, 3]) / self.num_samples
        ratios = counts / self.num_samples
        self.topk_invalid, self.topk_vocab, self.topk_guide, self.topk = ratios[:, 3], ratios[:, 2], ratios[:, 1], ratios[:, 0]
        return self
