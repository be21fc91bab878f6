"""Embedding-cache reader and device feeder: the step in front of the training path.

The reference trains from a memory-mapped cache file (embedding_cache.py:24-31 documents the format, :35-73 the 128-byte
header, :77-157 the section offsets) and forms each batch on the host: slice the embedding rows, gather every sample's M
target-noun tokenisations out of the R x C table with fancy indexing (:714-719), post-process (:827-895), then copy four
tensors to the device (:944-958).  Here the R x C token table and its padding mask are uploaded ONCE; a batch costs one
pinned-memory staging copy and one host->device transfer of [embeddings | noun ids | weights] on a side stream (double
buffered, so batch i+1 travels while step i computes), and the per-sample gather runs on the device.

File format (little endian): header | R null-separated UTF-8 nouns (first = '') | R x C token ids | R x C padding mask |
N x M noun ids | N x M weights | N x F unit embeddings; without targets only header + embeddings.

`write_cache` writes the same format (used for tests and synthetic workloads; hashes are zero and `embedder_strict` is off, so
the reference's `EmbeddingCache(..., strict_embedder=False)` opens the file - tests/test_cache.py checks exactly that).
"""
from __future__ import annotations

import dataclasses
import mmap
import os
import struct
from typing import Iterator, Optional, Sequence

import numpy as np
import torch

MAGIC = b'\xa9\xfdK\x14*\x9a\xb8\x13m\x157\xca\xe8+\xef\x82B\x19\xdbJ\xb8\x93\xb2&\xa0\x1a=\xe4\xadR\xb1\x99'   # embedding_cache.py:40
HEADER_STRUCT = struct.Struct('<32sB?????32s32sLLHHHLHHHH')                                                          # embedding_cache.py:42
assert HEADER_STRUCT.size == 128
VERSION = 1
INT_DTYPES = (torch.int8, torch.int16, torch.int32, torch.int64)                                                     # :49
BOOL_DTYPES = (torch.bool,)                                                                                          # :51
FLOAT_DTYPES = (torch.float16, torch.bfloat16, torch.float32, torch.float64)                                         # :53
_TORCH_OF_NP = {np.dtype(np.int8): torch.int8, np.dtype(np.int16): torch.int16, np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64,
                np.dtype(np.float16): torch.float16, np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}
_NP = {torch.int8: np.int8, torch.int16: np.int16, torch.int32: np.int32, torch.int64: np.int64, torch.bool: np.bool_,
       torch.float16: np.float16, torch.float32: np.float32, torch.float64: np.float64}


@dataclasses.dataclass(frozen=True)
class CacheHeader:
    """Fields in file order (embedding_cache.py:57-73)."""
    magic_bytes: bytes
    version: int
    use_targets: bool
    full_targets: bool
    default_weights: bool
    unit_weights: bool
    embedder_strict: bool
    embedder_hash: bytes
    target_config_hash: bytes
    target_nouns_num: int        # R
    target_nouns_size: int       # bytes of the noun strings
    target_dim: int              # C
    target_dtype_id: int
    target_mask_dtype_id: int
    embed_num: int               # N
    embed_targets_dim: int       # M
    embed_targets_dtype_id: int
    embed_dim: int               # F
    embed_dtype_id: int

    def pack(self) -> bytes:
        return HEADER_STRUCT.pack(*dataclasses.astuple(self))

    @staticmethod
    def unpack(raw: bytes) -> "CacheHeader":
        if len(raw) != HEADER_STRUCT.size:
            raise ValueError(f"Cache file too short for header: {len(raw)} bytes read but {HEADER_STRUCT.size} needed")
        return CacheHeader(*HEADER_STRUCT.unpack(raw))


@dataclasses.dataclass(frozen=True)
class CacheLayout:
    """Byte offsets of the sections (embedding_cache.py:115-157).  A cache without targets has zero-sized target sections in the
    reference's arithmetic only if its header says so; its writer zeroes R, C and M in that case."""
    target_dtype: torch.dtype
    mask_dtype: torch.dtype
    embed_targets_dtype: torch.dtype
    embed_dtype: torch.dtype
    target_nouns_offset: int
    target_offset: int
    target_mask_offset: int
    embed_targets_offset: int
    embed_target_weights_offset: int
    embed_offset: int
    embed_stride: int
    total_size: int

    @staticmethod
    def from_header(h: CacheHeader) -> "CacheLayout":
        td, md = INT_DTYPES[h.target_dtype_id], BOOL_DTYPES[h.target_mask_dtype_id]
        etd, ed = INT_DTYPES[h.embed_targets_dtype_id], FLOAT_DTYPES[h.embed_dtype_id]
        size = lambda dt: torch.tensor((), dtype=dt).element_size()  # noqa: E731
        nouns = HEADER_STRUCT.size
        tgt = nouns + h.target_nouns_size
        msk = tgt + h.target_nouns_num * h.target_dim * size(td)
        ids = msk + h.target_nouns_num * h.target_dim * size(md)
        wts = ids + h.embed_num * h.embed_targets_dim * size(etd)
        emb = wts + h.embed_num * h.embed_targets_dim * size(ed)
        stride = h.embed_dim * size(ed)
        return CacheLayout(td, md, etd, ed, nouns, tgt, msk, ids, wts, emb, stride, emb + h.embed_num * stride)


def write_cache(path: str, embeds: torch.Tensor, target_nouns: Optional[Sequence[str]] = None, target_token_ids: Optional[torch.Tensor] = None,
                target_mask: Optional[torch.Tensor] = None, embed_targets: Optional[torch.Tensor] = None,
                embed_target_weights: Optional[torch.Tensor] = None, unit_weights: bool = True) -> CacheHeader:
    """Write a cache file in the reference's format.  With targets: `target_nouns` are the R-1 real nouns (the empty noun of id 0
    and its fully padded tokenisation are prepended here, embedding_cache.py:28-29), `target_token_ids` / `target_mask` are
    (R-1) x C, `embed_targets` N x M noun ids in [0, R) with all zeros trailing, `embed_target_weights` N x M or None (uniform
    over the non-zero ids)."""
    N, F = embeds.shape
    use_targets = target_nouns is not None
    if use_targets:
        nouns = ('',) + tuple(target_nouns)
        if any('\x00' in n for n in nouns) or any(n == '' for n in nouns[1:]):
            raise ValueError("target nouns must be non-empty and free of null characters")
        R, C = len(nouns), target_token_ids.shape[1]
        assert target_token_ids.shape == (R - 1, C) and target_mask.shape == (R - 1, C) and target_mask.dtype == torch.bool
        assert embed_targets.shape[0] == N and int(embed_targets.min()) >= 0 and int(embed_targets.max()) < R
        assert bool((embed_targets[:, 0] != 0).all()), "the first target of every embedding must be a real noun (embedding_cache.py:30)"
        M = embed_targets.shape[1]
        noun_bytes = '\x00'.join(nouns).encode('utf-8')
        tok = torch.cat((torch.zeros(1, C, dtype=target_token_ids.dtype), target_token_ids))
        msk = torch.cat((torch.ones(1, C, dtype=torch.bool), target_mask))
        default_weights = embed_target_weights is None
        if default_weights:
            nz = (embed_targets != 0).to(embeds.dtype)
            embed_target_weights = nz / nz.sum(dim=1, keepdim=True)
        full_targets = bool((embed_targets != 0).all())
        header = CacheHeader(MAGIC, VERSION, True, full_targets, default_weights, unit_weights or default_weights, False, b'\x00' * 32, b'\x00' * 32, R,
                             len(noun_bytes), C, INT_DTYPES.index(tok.dtype), 0, N, M, INT_DTYPES.index(embed_targets.dtype), F, FLOAT_DTYPES.index(embeds.dtype))
        sections = (noun_bytes, tok.contiguous().numpy().tobytes(), msk.contiguous().numpy().tobytes(), embed_targets.contiguous().numpy().tobytes(),
                    embed_target_weights.to(embeds.dtype).contiguous().numpy().tobytes())
    else:
        header = CacheHeader(MAGIC, VERSION, False, False, False, False, False, b'\x00' * 32, b'\x00' * 32, 0, 0, 0, 0, 0, N, 0, 0, F, FLOAT_DTYPES.index(embeds.dtype))
        sections = ()
    with open(path, 'wb') as f:
        f.write(header.pack())
        for s in sections:
            f.write(s)
        f.write(embeds.contiguous().numpy().tobytes())
    assert os.path.getsize(path) == CacheLayout.from_header(header).total_size
    return header


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class EmbeddingCacheReader:
    """Memory-mapped cache file -> batches on `device` (a CUDA device, or 'cpu' for host-side use and tests).

    get_samples(start, stop, use_weights) returns the reference's 5-tuple (embedding_cache.py:699-723):
    (embed B x F, target_ids B x M, target B x M x C, mask B x M x C, weight B x M or None), on `device`.
    batches(batch_size, training, epoch_index_offset) yields the reference Dataset's batches (:827-895): (embed, target, mask, weight),
    single-target caches squeezed to B x C, trailing all-padded token columns trimmed, training mode dropping the incomplete batch
    and wrapping around the end of the file by the epoch offset."""

    def __init__(self, path: str, device="cuda", embed_dim: Optional[int] = None, use_targets: Optional[bool] = None):
        self.path = os.path.abspath(path)
        self.device = torch.device(device)
        with open(self.path, 'rb') as f:
            self.header = CacheHeader.unpack(f.read(HEADER_STRUCT.size))
            h = self.header
            if h.magic_bytes != MAGIC:
                raise ValueError("Cache file has invalid magic bytes (unfinished or foreign file)")        # embedding_cache.py:494-495
            if h.version > VERSION or h.version < 1:
                raise ValueError(f"Cache file version is unsupported: {h.version} vs supported {VERSION}")
            self.use_targets = h.use_targets if use_targets is None else bool(use_targets)
            if self.use_targets and not h.use_targets:
                raise ValueError("Embedding cache reader requires targets but the cache file has none")
            self.layout = CacheLayout.from_header(h)
            if self.use_targets:
                raw = f.read(h.target_nouns_size)
                self.target_nouns = tuple(raw.decode('utf-8').split('\x00'))
                if len(self.target_nouns) != h.target_nouns_num or self.target_nouns[0] != '':
                    raise ValueError("Cache file target nouns are inconsistent with its header")
            else:
                self.target_nouns = None
            f.seek(0, os.SEEK_END)
            if f.tell() != self.layout.total_size:
                raise ValueError(f"Cache file has an unexpected actual size: {f.tell()} vs {self.layout.total_size}")
        if h.embed_num < 1:
            raise ValueError(f"Cache file must have a positive number of embeddings: {h.embed_num}")
        if embed_dim is not None and h.embed_dim != embed_dim:
            raise ValueError(f"Cache file has embedding dimension mismatch: {h.embed_dim} vs {embed_dim}")
        if self.layout.embed_dtype == torch.bfloat16:
            raise ValueError("bfloat16 embedding caches are not supported by this reader")
        self._file = open(self.path, 'rb')
        self._mmap = mmap.mmap(self._file.fileno(), length=0, access=mmap.ACCESS_READ)
        L = self.layout
        view = lambda off, dt, shape: np.frombuffer(self._mmap, dtype=_NP[dt], count=int(np.prod(shape)), offset=off).reshape(shape)  # noqa: E731
        self._embed = view(L.embed_offset, L.embed_dtype, (h.embed_num, h.embed_dim))
        if self.use_targets:
            self._embed_targets = view(L.embed_targets_offset, L.embed_targets_dtype, (h.embed_num, h.embed_targets_dim))
            self._weights = view(L.embed_target_weights_offset, L.embed_dtype, (h.embed_num, h.embed_targets_dim))
            # the noun table lives on the device for the lifetime of the reader: R x C ids and padding
            self.target_token_ids = torch.from_numpy(view(L.target_offset, L.target_dtype, (h.target_nouns_num, h.target_dim)).copy()).to(self.device)
            self.target_mask = torch.from_numpy(view(L.target_mask_offset, L.mask_dtype, (h.target_nouns_num, h.target_dim)).copy()).to(self.device)
        self._pinned = {}
        self._slot_events = {}
        self._copy_stream = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None

    def __len__(self) -> int:
        return self.header.embed_num

    def close(self) -> None:
        for name in ("_embed", "_embed_targets", "_weights"):
            if hasattr(self, name):
                delattr(self, name)
        if getattr(self, "_mmap", None) is not None:
            self._mmap.close()
            self._mmap = None
        if getattr(self, "_file", None) is not None:
            self._file.close()
            self._file = None

    def __enter__(self) -> "EmbeddingCacheReader":
        return self

    def __exit__(self, *exc) -> bool:
        self.close()
        return False

    # ------------------------------------------------------------------------------------------------------------
    def _staged(self, slot: int, start: int, stop: int, use_weights: bool):
        """Rows [start, stop) -> pinned staging buffers of `slot` -> device, on the copy stream when there is one.  Returns the device
        tensors and the CUDA event that marks the end of their transfer (None on the CPU)."""
        n = stop - start
        h = self.header
        parts = [("embed", self._embed, h.embed_dim)]
        if self.use_targets:
            parts.append(("ids", self._embed_targets, h.embed_targets_dim))
            if use_weights:
                parts.append(("weight", self._weights, h.embed_targets_dim))
        cuda = self.device.type == "cuda"
        prev = self._slot_events.get(slot)
        if prev is not None:
            prev.synchronize()                                  # the slot's previous transfer has left the pinned buffers
        out = {}
        ctx = torch.cuda.stream(self._copy_stream) if cuda else _NullCtx()
        with ctx:
            for name, src, width in parts:
                key = (slot, name)
                buf = self._pinned.get(key)
                if buf is None or buf.shape[0] < n:
                    buf = torch.empty((max(n, 1), width), dtype=_TORCH_OF_NP[src.dtype])
                    if cuda:
                        buf = buf.pin_memory()
                    self._pinned[key] = buf
                np.copyto(buf[:n].numpy(), src[start:stop])
                out[name] = buf[:n].to(self.device, non_blocking=True) if cuda else buf[:n].clone()
            ev = None
            if cuda:
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
                self._slot_events[slot] = ev
        return out, ev

    def _finish(self, staged, use_weights: bool):
        """Make the current stream wait for the transfer, then gather the token rows of every sample's nouns on the device."""
        dev, ev = staged
        if ev is not None:
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            for t in dev.values():
                t.record_stream(cur)
        embed = dev["embed"]
        if not self.use_targets:
            return embed, None, None, None, None
        ids = dev["ids"]
        idx = ids.to(torch.int64)
        target = self.target_token_ids[idx]                    # B x M x C, gathered on the device (embedding_cache.py:716)
        mask = self.target_mask[idx]
        return embed, ids, target, mask, (dev.get("weight") if use_weights else None)

    def get_samples(self, start: int, stop: int, use_weights: bool = True):
        if start < 0 or stop < 0:
            raise IndexError("Negative indices are not supported")
        stop = min(stop, self.header.embed_num)
        if stop - start <= 0:
            start = stop = 0
        return self._finish(self._staged(0, start, stop, use_weights), use_weights)

    # ------------------------------------------------------------------------------------------------------------
    def num_batches(self, batch_size: int, training: bool) -> int:
        full, rest = divmod(self.header.embed_num, batch_size)
        return full if training or rest == 0 else full + 1     # embedding_cache.py:776-787

    def _batch_rows(self, index: int, batch_size: int, training: bool, epoch_index_offset: int):
        N = self.header.embed_num
        if epoch_index_offset == 0 or not training:           # :832-834
            start = index * batch_size
            return [(start, min(start + batch_size, N))]
        start = (index * batch_size + epoch_index_offset) % N  # :836-841
        stop = (start + batch_size - 1) % N + 1
        return [(start, stop)] if start < stop else [(start, N), (0, stop)]

    def _post(self, embed, ids, target, mask, weight, fixed_token_length: bool):
        """Dataset.__getitem__ post-processing (embedding_cache.py:843-893) for the nominal data configuration of the file:
        multi_target iff M > 1, multi_first False, unit weights as stored."""
        if ids is None:
            return embed, None, None, None
        h = self.header
        if h.embed_targets_dim > 1:
            if target.shape[1] > 1:                                # drop trailing target slots nobody in the batch uses (:860-867)
                used = (ids != 0 if weight is None else weight != 0).any(dim=0)
                if not bool(used.all()):
                    k = int((~used).to(torch.int8).argmax())
                    target, mask = target[:, :k], mask[:, :k]
                    weight = None if weight is None else weight[:, :k]
        else:
            target, mask = target[:, 0], mask[:, 0]
            weight = None if weight is None else weight[:, 0]
        if not fixed_token_length:                                 # drop trailing token columns that are padding everywhere (:887-891)
            col = mask.flatten(0, -2).all(dim=0)
            if bool(col.any()):
                k = int(col.to(torch.int8).argmax())
                target, mask = target[..., :k], mask[..., :k]
        return embed, target, mask, weight

    def batches(self, batch_size: int, training: bool = False, epoch_index_offset: int = 0, use_weights: Optional[bool] = None,
                fixed_token_length: bool = False) -> Iterator[tuple]:
        """Yield (embed, target, mask, weight) batches on the device, the next batch's transfer overlapping the caller's work."""
        h = self.header
        if batch_size < 1 or batch_size > h.embed_num:
            raise ValueError(f"Batch size must be in [1, {h.embed_num}]: {batch_size}")
        if use_weights is None:
            use_weights = self.use_targets and not (h.default_weights and h.full_targets)      # :795
        n = self.num_batches(batch_size, training)

        def load(i):      # start the transfers of batch i (two pieces when it wraps around the end of the file); slots 1..4, 0 is get_samples'
            return [self._staged(1 + 2 * (i & 1) + k, a, b, use_weights) for k, (a, b) in enumerate(self._batch_rows(i, batch_size, training, epoch_index_offset))]

        def finish(staged):
            pieces = [self._finish(st, use_weights) for st in staged]
            if len(pieces) == 1:
                return pieces[0]
            return tuple(None if p[0] is None else torch.cat(p, dim=0) for p in zip(*pieces))

        nxt = load(0) if n > 0 else None
        for i in range(n):
            cur = nxt
            nxt = load(i + 1) if i + 1 < n else None          # batch i + 1 travels while the caller works on batch i
            yield self._post(*finish(cur), fixed_token_length)
