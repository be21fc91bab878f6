"""Target-noun id formats on the caller's side of the decoder: the tokenizer-independent half of
`Embedder.create_target_config` / `tokenize_target` / `detokenize_target` (embedders.py:169-254, :331-385, :387-406).

The reference turns a tokenizer's raw output (ids padded to the longest text, attention mask, optional start token, end token, pad
token) into the decoder's target format - optional start / end tokens, ids renumbered densely over the tokens the noun vocabulary
actually uses (pad = 0, end = 0, start = 1), optional fixed length, padding mask - and back before the ids are handed to the
tokenizer's own `detokenize`.  Those steps are pure id arithmetic; here they are batched tensor operations that run on whatever device
the ids live on (the reference walks Python sets and lists per batch), so decoded ids can be mapped back on the GPU before the one
device->host copy of a serving step.  Turning raw ids into strings needs the CLIP tokenizer's vocabulary, which is not vendored with
the reference and not available offline - that last step stays with the reference's embedder.

Parity: tests/test_targets.py drives the unmodified reference `Embedder` (with a toy word-piece tokenizer) through every combination of
the format switches and requires identical configurations, ids and masks.
"""
from __future__ import annotations

import dataclasses
from typing import Optional

import torch


@dataclasses.dataclass(frozen=True)
class TokenizerIds:
    """What the formats need to know about a tokenizer (Embedder.__init__, embedders.py:96-130)."""
    vocab_size: int
    start_token_id: Optional[int]      # None: the tokenizer emits no start token
    end_token_id: int
    pad_token_id: int
    context_length: int
    token_dtype: torch.dtype = torch.int64


@dataclasses.dataclass(frozen=True)
class TargetFormat:
    """Field for field the reference's TargetConfig (embedders.py:42-65) minus mask_dtype (always bool)."""
    vocab_size: int
    token_dtype: torch.dtype
    start_token_id: Optional[int]
    end_token_id: Optional[int]
    pad_token_id: int
    compact_ids: bool
    compact_map: Optional[torch.Tensor]     # [tokenizer vocab] -> compact id, -1 where unused
    compact_unmap: Optional[torch.Tensor]   # [vocab_size] -> tokenizer id (-1 for a start token the tokenizer does not have)
    fixed_token_length: bool
    token_length: int
    use_masks: bool


def make_target_format(ids: torch.Tensor, attention_mask: torch.Tensor, tok: TokenizerIds, *, with_start_token: bool, with_end_token: bool,
                       compact_ids: bool, fixed_token_length: bool, auto_fixed_token_length: bool, use_masks: bool) -> TargetFormat:
    """create_target_config (embedders.py:169-254) from the raw tokenization of ALL target nouns: ids / attention_mask are [N, T], padded
    to the longest noun (T counts the tokenizer's own start token, if it has one, and the end token)."""
    if ids.ndim != 2 or ids.shape != attention_mask.shape or ids.shape[0] < 1:
        raise ValueError("ids and attention_mask must be non-empty [N, T] tensors of the same shape")
    max_tokens = int(attention_mask.to(torch.bool).sum(dim=1).max())
    if not with_end_token:
        max_tokens -= 1
    if tok.start_token_id is None:
        if with_start_token:
            max_tokens += 1
    elif not with_start_token:
        max_tokens -= 1
    if compact_ids:
        used = torch.zeros(tok.vocab_size, dtype=torch.bool, device=ids.device)
        used[ids.reshape(-1).long()] = True
        if not bool(used[tok.end_token_id]):
            raise KeyError(tok.end_token_id)              # the reference's set.remove raises when no text carries an end token
        used[tok.end_token_id] = False
        used[tok.pad_token_id] = False
        if tok.start_token_id is not None:
            if not bool(used[tok.start_token_id]):
                raise KeyError(tok.start_token_id)
            used[tok.start_token_id] = False
        content = used.nonzero().reshape(-1).to(tok.token_dtype).cpu()      # ascending = sorted(token_id_set)
        special = [tok.pad_token_id]
        if with_start_token:
            special.append(tok.start_token_id if tok.start_token_id is not None else -1)
        num_special = len(special)
        compact_unmap = torch.cat((torch.tensor(special, dtype=tok.token_dtype), content))
        vocab_size = compact_unmap.numel()
        compact_map = torch.full((tok.vocab_size,), -1, dtype=tok.token_dtype)
        compact_map[content.long()] = torch.arange(num_special, vocab_size, dtype=tok.token_dtype)
        compact_map[tok.pad_token_id] = 0
        compact_map[tok.end_token_id] = 0
        if tok.start_token_id is not None and with_start_token:
            compact_map[tok.start_token_id] = 1
        start_id, end_id, pad_id = (1 if with_start_token else None), (0 if with_end_token else None), 0
    else:
        vocab_size, compact_map, compact_unmap = tok.vocab_size, None, None
        start_id = tok.start_token_id if with_start_token else None
        end_id = tok.end_token_id if with_end_token else None
        pad_id = tok.pad_token_id
    token_length = max_tokens if (not fixed_token_length or auto_fixed_token_length) else tok.context_length
    return TargetFormat(vocab_size=vocab_size, token_dtype=tok.token_dtype, start_token_id=start_id, end_token_id=end_id, pad_token_id=pad_id,
                        compact_ids=compact_ids, compact_map=compact_map, compact_unmap=compact_unmap, fixed_token_length=fixed_token_length,
                        token_length=token_length, use_masks=use_masks)


def encode_targets(ids: torch.Tensor, attention_mask: torch.Tensor, tok: TokenizerIds, fmt: TargetFormat) -> tuple[torch.Tensor, Optional[torch.Tensor]]:
    """tokenize_target after the tokenizer has run (embedders.py:339-366): raw [B, T] ids / attention mask (padded to the batch's longest
    text) -> (target ids, padding mask or None) in the decoder's format."""
    T = ids.shape[1]
    skip_start = 1 if tok.start_token_id is not None and fmt.start_token_id is None else 0
    skip_end = T - 1 if fmt.end_token_id is None else T
    out = ids[:, skip_start:skip_end].clone()
    mask = torch.logical_not(attention_mask[:, skip_start:skip_end].to(torch.bool)) if fmt.use_masks else None
    if fmt.compact_ids:
        if fmt.end_token_id is None and mask is not None:
            mask = mask | (out == tok.end_token_id)
        out = fmt.compact_map.to(out.device)[out.long()]          # maps the end token to the pad id when the format has no end token
        if tok.start_token_id is None and fmt.start_token_id is not None:
            out = torch.cat((out.new_ones((out.shape[0], 1)), out), dim=1)
            if mask is not None:
                mask = torch.cat((mask.new_zeros((mask.shape[0], 1)), mask), dim=1)
    elif fmt.end_token_id is None:
        is_end = out == tok.end_token_id
        out[is_end] = fmt.pad_token_id
        if mask is not None:
            mask = mask | is_end
    if fmt.fixed_token_length:
        L = out.shape[1]
        if L > fmt.token_length:
            raise ValueError(f"Sequence length {L} is larger than the configured target tokenization fixed length {fmt.token_length}")
        if L < fmt.token_length:
            out = torch.cat((out, out.new_full((out.shape[0], fmt.token_length - L), fmt.pad_token_id)), dim=1)
            if mask is not None:
                mask = torch.cat((mask, mask.new_ones((mask.shape[0], fmt.token_length - L))), dim=1)
    return out, mask


def decode_targets(token_ids: torch.Tensor, tok: TokenizerIds, fmt: TargetFormat) -> torch.Tensor:
    """detokenize_target up to the tokenizer's own detokenize (embedders.py:395-400): target ids of any leading shape [..., S] -> tokenizer ids."""
    if fmt.compact_ids:
        if tok.start_token_id is None and fmt.start_token_id is not None:
            token_ids = token_ids[..., 1:]
        token_ids = fmt.compact_unmap.to(token_ids.device)[token_ids.long()]
    return token_ids
