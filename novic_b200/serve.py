"""Pipelined batch inference around PrefixedIterDecoder.generate / generate_beam: the loop of infer.py (NOVICModel.embed ->
GenerationTask.generate -> GenerationTask.update, infer.py:240-251, :556-644) with the transfers taken off the critical path.

The reference copies a batch to the device, decodes it, then reads the results back before it touches the next batch.  Here the
host->device copy of batch i+1 runs on a side stream while batch i decodes, and the device->host copy of batch i's ids / padding /
scores (into pinned buffers) runs while batch i+1 decodes; results are handed out one batch late, in order.  Decoding itself is
untouched - the same generate / generate_beam calls, so every output is identical to calling them batch by batch.
"""
from __future__ import annotations

from typing import Callable, Iterable, Iterator, Optional

import torch


class GenerationPipeline:
    """run(host_batches) yields (target, target_padding, target_score) per batch as CPU tensors, shaped like GenerationTask.generate
    returns them (infer.py:556-611): [B, K, T] ids, [B, K, T] padding, [B, K] scores (K = 1 for greedy).

    method: 'greedy' | 'beam'; the remaining arguments are passed through to generate / generate_beam.
    post:   optional callable (tok, pad, score, T) -> (tok, pad, score, T) applied to the device results before they leave the GPU
            (e.g. a closure over novic_b200.dist.gather_generation_async); T is None or a one-element device tensor holding the
            early-exit length that tok / pad still have to be cut to;
    emit:   False = this process does not read results back (non-root ranks of a sharded job); None is yielded instead."""

    def __init__(self, model, method: str = "greedy", topk: int = 1, temperature: float = 1.0, length_alpha: float = 0.0,
                 guide_targets: Optional[torch.Tensor] = None, guide_renorm: bool = False, vocab_targets: Optional[torch.Tensor] = None,
                 vocab_per_token: bool = False, vocab_scaler: float = 0.0, post: Optional[Callable] = None, emit: bool = True):
        if method not in ("greedy", "beam"):
            raise ValueError(f"Unsupported generation method: {method}")
        self.model, self.method, self.topk = model, method, int(topk)
        self.temperature, self.length_alpha = float(temperature), float(length_alpha)
        self.guide_targets, self.guide_renorm = guide_targets, bool(guide_renorm)
        self.vocab_targets, self.vocab_per_token, self.vocab_scaler = vocab_targets, bool(vocab_per_token), float(vocab_scaler)
        self.post, self.emit = post, emit
        self.device = next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("GenerationPipeline needs the model on a CUDA device: novic_b200 has no CPU path")
        self._copy = torch.cuda.Stream(self.device)
        self._pinned = [None, None]

    def _decode(self, embed):
        """-> (tok [B, K, T'], pad, score [B, K], T): T is None when tok / pad are already cut to the early-exit length, else a
        one-element device tensor holding it (greedy: nothing here waits for the GPU, the next batch can be enqueued right away)."""
        m = self.model
        T = None
        if self.method == "greedy":
            from .decoder import MAX_SEQS_PER_CALL
            if embed.shape[0] <= MAX_SEQS_PER_CALL:
                tok, pad, score, T = m.generate_async(embed, self.temperature, self.length_alpha, self.guide_targets, self.guide_renorm)
            else:
                tok, pad, _, _, _, score = m.generate(embed, False, True, self.temperature, self.length_alpha, None, self.guide_targets, self.guide_renorm)
            tok, pad, score = tok.unsqueeze(1), pad.unsqueeze(1), score.unsqueeze(1)
        else:
            tok, pad, score = m.generate_beam(embed, self.topk, self.temperature, self.length_alpha, self.vocab_targets, self.vocab_per_token,
                                              self.vocab_scaler, self.guide_targets, self.guide_renorm)
        if self.post is not None:
            tok, pad, score, T = self.post(tok, pad, score, T)
        return tok, pad, score, T

    @staticmethod
    def _cut(outs):
        tok, pad, score, T = outs
        if T is not None:
            t = int(T.reshape(-1)[0])
            tok, pad = tok[:, :, :t], pad[:, :, :t]
        return tok.clone(), pad.clone(), score.clone()

    def _upload(self, host: torch.Tensor):
        with torch.cuda.stream(self._copy):
            dev = host.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy)
        return dev, ev

    def _download(self, slot: int, res, after: torch.cuda.Event):
        """Start the device->host copies of one batch's results on the copy stream; returns (pinned views, completion event)."""
        self._copy.wait_event(after)
        outs = []
        with torch.cuda.stream(self._copy):
            bufs = self._pinned[slot] or [None, None, None, None]
            bufs = (bufs + [None] * 4)[:4]
            for k, t in enumerate(res):
                if t is None:
                    outs.append(None)
                    continue
                t = t.contiguous()
                if bufs[k] is None or bufs[k].numel() < t.numel() or bufs[k].dtype != t.dtype:
                    bufs[k] = torch.empty(t.numel(), dtype=t.dtype).pin_memory()
                view = bufs[k][: t.numel()].view(t.shape)
                view.copy_(t, non_blocking=True)
                t.record_stream(self._copy)
                outs.append(view)
            self._pinned[slot] = bufs
            ev = torch.cuda.Event()
            ev.record(self._copy)
        return outs, ev

    def run(self, host_batches: Iterable[torch.Tensor]) -> Iterator:
        cur_stream = torch.cuda.current_stream(self.device)
        it = iter(host_batches)
        try:
            nxt = self._upload(next(it))
        except StopIteration:
            return
        pending = None           # (pinned views, event) of the previous batch's results
        i = 0
        # inference mode is entered around the library work only and left before every yield: a generator that yields inside
        # `with torch.inference_mode()` would leak the thread-local mode into the consumer's loop body
        while nxt is not None:
            dev, ev = nxt
            try:
                nxt = self._upload(next(it))          # batch i + 1 travels while batch i decodes
            except StopIteration:
                nxt = None
            ready = None
            with torch.inference_mode():
                cur_stream.wait_event(ev)
                dev.record_stream(cur_stream)
                res = self._decode(dev)
                done = torch.cuda.Event()
                done.record(cur_stream)
                if pending is not None:                   # hand out batch i - 1 (its copies finished during this decode)
                    outs, dev_ev = pending
                    dev_ev.synchronize()
                    ready = self._cut(outs)
                pending = self._download(i & 1, res, done) if self.emit else None
            if ready is not None:
                yield ready
            if not self.emit:
                yield None
            i += 1
        if pending is not None:
            outs, dev_ev = pending
            dev_ev.synchronize()
            with torch.inference_mode():
                last = self._cut(outs)
            yield last
