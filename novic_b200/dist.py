"""Multi-GPU plumbing for the inference path: one process per GPU, contiguous shards of the embedding batch,
no collective inside the decode loop, one final gather of token ids / padding / scores.

Only torch.distributed is used (NCCL on GPUs; the same code runs over gloo on CPU tensors for the tests).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


def shard_bounds(n: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous split of n items: the first (n % world_size) ranks get one extra item."""
    base, extra = divmod(n, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def _pad_cols(t: torch.Tensor, T: int, dim: int, value) -> torch.Tensor:
    if t.shape[dim] == T:
        return t
    shape = list(t.shape)
    shape[dim] = T - t.shape[dim]
    return torch.cat((t, torch.full(shape, value, dtype=t.dtype, device=t.device)), dim=dim)


def gather_generation(tok: torch.Tensor, pad: torch.Tensor, score: torch.Tensor, total: int,
                      group: Optional[dist.ProcessGroup] = None) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """All-gather the per-rank results of generate / generate_beam into the full batch, in rank order.

    tok / pad are [n_local, (H,) T_local], score is [n_local(, H)].  Ranks may have stopped at different T (early
    exit is per shard): columns are padded to the global maximum with (id 0, padding True) - exactly what the
    reference returns for samples that finished before the longest one.  Shards may differ in size by one row.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    tdim = tok.ndim - 1
    t_max = torch.tensor([tok.shape[tdim]], dtype=torch.int64, device=tok.device)
    dist.all_reduce(t_max, op=dist.ReduceOp.MAX, group=group)
    T = int(t_max.item())
    tok = _pad_cols(tok, T, tdim, 0)
    pad = _pad_cols(pad, T, tdim, True)
    n_max = -(-total // world)
    n_local = tok.shape[0]
    assert (n_local,) == (shard_bounds(total, world, rank)[1] - shard_bounds(total, world, rank)[0],)

    def gather(t: torch.Tensor, fill) -> torch.Tensor:
        t = _pad_cols(t.contiguous(), n_max, 0, fill)      # equal-sized contributions
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t, group=group)
        sizes = [shard_bounds(total, world, r) for r in range(world)]
        return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(outs, sizes)], dim=0)

    pad_u8 = pad.to(torch.uint8)
    return gather(tok, 0), gather(pad_u8, 1).to(torch.bool), gather(score, float("-inf"))


def generate_sharded(model, embed_full: torch.Tensor, method: str = "greedy", topk: int = 1, temperature: float = 1.0,
                     length_alpha: float = 0.0, group: Optional[dist.ProcessGroup] = None):
    """Decode this rank's contiguous shard of `embed_full` (every rank holds or can index the full batch) and gather.
    Returns (tok, pad, score) for the whole batch on every rank, shaped like GenerationTask.generate (infer.py:556-611):
    [B, K, T], [B, K, T], [B, K]."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_bounds(embed_full.shape[0], world, rank)
    local = embed_full[lo:hi]
    if method == "greedy":
        tok, pad, _, _, _, score = model.generate(local, False, True, temperature, length_alpha, None, None, False)
        tok, pad, score = tok.unsqueeze(1), pad.unsqueeze(1), score.unsqueeze(1)
    elif method == "beam":
        tok, pad, score = model.generate_beam(local, topk, temperature, length_alpha, None, False, 0.0, None, False)
    else:
        raise ValueError(f"Unsupported generation method: {method}")
    return gather_generation(tok, pad, score, embed_full.shape[0], group)
