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


def gather_generation_async(tok: torch.Tensor, pad: torch.Tensor, score: torch.Tensor, total: int, gen_len: int,
                            T_local=None, group: Optional[dist.ProcessGroup] = None) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """All-gather the per-rank results of generate / generate_beam with ONE collective and no host synchronisation.

    tok / pad are [n_local, K, T_have] with T_have <= gen_len columns present, score is [n_local, K]; T_local is this shard's early-exit
    length as a python int or a device int32/int64 tensor (default: T_have).  Returns (tok [total, K, gen_len], pad [total, K, gen_len]
    bool, score [total, K], T int64 device scalar = global maximum early-exit length): the caller cuts the columns to T once it is on
    the host.  Columns at or beyond a rank's own T hold id 0 / padding True - exactly what the reference returns for samples that finished
    before the longest one.  Shards may differ in size by one row.  The payload is one byte buffer of four 8-byte aligned sections:
    [int64 ids | fp32 scores | u8 padding | int64 T_local], each a typed view - a handful of copies to pack, none to unpack."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_local, K, T_have = tok.shape
    G = int(gen_len)
    lo, hi = shard_bounds(total, world, rank)
    assert n_local == hi - lo and T_have <= G
    n_max = -(-total // world)
    rows = n_max * K
    o_ids, o_score = 0, rows * G * 8
    o_pad = o_score + rows * 4
    o_hdr = (o_pad + rows * G + 7) // 8 * 8
    nbytes = o_hdr + 8
    dev = tok.device
    full = n_local == n_max and T_have == G                  # every byte of the payload gets written below: no need to clear it first
    buf = torch.empty(nbytes, dtype=torch.uint8, device=dev) if full else torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    buf[o_ids:o_score].view(torch.int64).view(n_max, K, G)[:n_local, :, :T_have] = tok
    buf[o_score:o_pad].view(torch.float32).view(n_max, K)[:n_local] = score
    pd = buf[o_pad:o_pad + rows * G].view(n_max, K, G)
    if not full:
        pd[:n_local, :, T_have:] = 1
    pd[:n_local, :, :T_have] = pad.view(torch.uint8) if pad.dtype == torch.bool else pad
    hdr = buf[o_hdr:].view(torch.int64)
    if T_local is None:
        hdr.fill_(T_have)
    elif isinstance(T_local, torch.Tensor):
        hdr.copy_(T_local.reshape(1))
    else:
        hdr.fill_(int(T_local))
    out = torch.empty((world, nbytes), dtype=torch.uint8, device=dev)
    dist.all_gather(list(out.unbind(0)), buf, group=group)
    T = out[:, o_hdr:].contiguous().view(torch.int64).max()
    sizes = [shard_bounds(total, world, r) for r in range(world)]
    even = all(hi_ - lo_ == n_max for lo_, hi_ in sizes)
    ids_all = out[:, o_ids:o_score].contiguous().view(torch.int64).view(world, n_max, K, G)
    sc_all = out[:, o_score:o_pad].contiguous().view(torch.float32).view(world, n_max, K)
    pd_all = out[:, o_pad:o_pad + rows * G].view(world, n_max, K, G)
    if even:
        toks, scores, pads = ids_all.flatten(0, 1), sc_all.flatten(0, 1), pd_all.flatten(0, 1)
    else:
        toks = torch.cat([ids_all[r, : hi_ - lo_] for r, (lo_, hi_) in enumerate(sizes)])
        scores = torch.cat([sc_all[r, : hi_ - lo_] for r, (lo_, hi_) in enumerate(sizes)])
        pads = torch.cat([pd_all[r, : hi_ - lo_] for r, (lo_, hi_) in enumerate(sizes)])
    return toks, pads.to(torch.bool), scores, T


def gather_generation(tok: torch.Tensor, pad: torch.Tensor, score: torch.Tensor, total: int,
                      group: Optional[dist.ProcessGroup] = None, gen_len: Optional[int] = None) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """All-gather the per-rank results of generate / generate_beam into the full batch, in rank order, cut to the global early-exit
    length (one collective, one host synchronisation to learn that length).  tok / pad are [n_local, K, T_local], score [n_local, K];
    ranks may have stopped at different T (early exit is per shard)."""
    G = int(gen_len) if gen_len is not None else None
    if G is None:  # callers that do not know the model's G: agree on the widest T first (one extra tiny collective)
        t_max = torch.tensor([tok.shape[2]], dtype=torch.int64, device=tok.device)
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX, group=group)
        G = int(t_max.item())
    toks, pads, scores, T = gather_generation_async(tok, pad, score, total, G, None, group)
    T = int(T.item())                                                          # the one host sync of the gather
    return toks[:, :, :T], pads[:, :, :T], scores


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
    return gather_generation(tok, pad, score, embed_full.shape[0], group, gen_len=model.target_config.token_length - 1 if hasattr(model, 'target_config') else None)


# ----------------------------------------------------------------------------------------------------------------
# Data-parallel training: gradients of the 40 parameter tensors are summed over ranks as one flat
# bucket (12.73 M fp32 = 50.9 MB, a single NCCL all-reduce over NVLink), together with the two loss scalars.
# ----------------------------------------------------------------------------------------------------------------
class GradBucket:
    """The flat fp32 buffer the training step's gradients are views of (novic_b200/training.py allocates one per step; autograd hands
    the views to `.grad` without copying).  `spare` trailing floats carry the loss scalars through the same collective."""
    SPARE = 8
    in_place_reductions = 0      # how many all-reduces ran directly on a bucket (diagnostics: tools/bench_train.py prints it)

    def __init__(self, flat: torch.Tensor, views):
        self.flat = flat
        self.total = sum(v.numel() for v in views)
        self.ptrs = {v.data_ptr(): v.numel() for v in views}

    def in_parameter_order(self, grads) -> bool:
        """True when `grads` lie in the buffer in the order given (the fused optimizer addresses parameters and gradients by one index)."""
        ptrs = [g.data_ptr() for g in grads]
        return all(a < b for a, b in zip(ptrs, ptrs[1:])) and (not ptrs or ptrs[0] == self.flat.data_ptr())

    def covers(self, grads) -> bool:
        """True when `grads` are exactly this bucket's views (same memory, every element accounted for once)."""
        if len(grads) != len(self.ptrs) or sum(g.numel() for g in grads) != self.total:
            return False
        seen = set()
        for g in grads:
            ptr = g.data_ptr()
            if g.dtype != torch.float32 or not g.is_contiguous() or self.ptrs.get(ptr) != g.numel() or ptr in seen:
                return False
            seen.add(ptr)
        return True


def allreduce_gradients(params, extra: Optional[torch.Tensor] = None, group: Optional[dist.ProcessGroup] = None,
                        bucket: Optional[GradBucket] = None) -> Optional[torch.Tensor]:
    """Sum `.grad` of `params` (and the optional small tensor `extra`, e.g. [loss_sum, loss_basis]) over all ranks, in
    place.  One flat bucket -> one collective.  When the gradients already live in `bucket` (the training step's own flat buffer)
    the collective runs on that buffer directly: no gather / scatter copies."""
    grads = [p.grad for p in params if p.grad is not None]
    if bucket is not None and bucket.covers(grads) and (extra is None or extra.numel() <= GradBucket.SPARE):
        flat = bucket.flat
        n = 0 if extra is None else extra.numel()
        tail = flat[bucket.total:bucket.total + GradBucket.SPARE]
        tail.zero_()
        if n:
            tail[:n].copy_(extra.reshape(-1).to(torch.float32))
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        GradBucket.in_place_reductions += 1
        return tail[:n].clone().view_as(extra).to(extra.dtype) if n else None
    pieces = grads + ([extra] if extra is not None else [])
    flat = torch.cat([t.reshape(-1).to(torch.float32) for t in pieces])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for t in grads:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n
    if extra is not None:
        return flat[off:off + extra.numel()].view_as(extra).to(extra.dtype)
    return None


_COMM = {}   # per device: (side stream, split event) of the overlapped gradient all-reduce


def _comm_objects(dev: torch.device):
    key = (dev.type, dev.index)
    if key not in _COMM:
        stream = torch.cuda.Stream(dev)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))      # creates the underlying cudaEvent_t (torch makes it lazily)
        _COMM[key] = (stream, ev)
    return _COMM[key]


def _flatten_multi(model, target, mask, weight):
    """B x M x C (or M x B x C with multi_first) -> A x C with the sequences of one embedding adjacent, as forward() does."""
    if target.ndim != 3:
        return target, mask, weight, 1
    dc = model.data_config
    if bool(getattr(dc, 'multi_target', False) and getattr(dc, 'multi_first', False)):
        target = target.transpose(0, 1)
        mask = None if mask is None else mask.transpose(0, 1)
        weight = None if weight is None else weight.transpose(0, 1)
    B, M = target.shape[:2]
    return (target.reshape(B * M, -1), None if mask is None else mask.reshape(B * M, -1), None if weight is None else weight.reshape(B * M), M)


def _train_step_fused(model, optimizer, embed, target, mask, weight, noise, gradient_clip, group):
    """The whole step on the device without a host synchronisation: noise -> forward + backward (one library call, CUDA graphs) ->
    all-reduce of the flat gradient bucket in two parts, the first one (layers >= split, final early in the backward pass) overlapping
    the rest of the backward pass on a side stream -> global-norm clip + AdamW (three kernels, optim.FusedAdamW)."""
    from . import training
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    if noise is not None:
        noise.stream_offset = rank
        embed = noise(embed)
    model.dropout_seed_offset = rank * 0x51ED2701      # per-rank dropout masks under one torch.manual_seed
    tgt, pad, w, M = _flatten_multi(model, target, mask, weight)
    pad_u8 = None if pad is None else pad.contiguous().view(torch.uint8)
    dev = embed.device
    L = len(model.transformer.layers)
    split = max(1, L // 3) if (world > 1 and L >= 2) else -1
    side, ev = _comm_objects(dev) if split > 0 else (None, None)
    loss, correct, pad_out, bucket = training.fwd_bwd(model, embed.contiguous(), tgt.contiguous(), pad_u8, None if w is None else w.contiguous(), M,
                                                     split_layer=split, split_event=ev)
    flat = bucket.flat
    stats = flat[bucket.total:bucket.total + 3]
    stats[:2].copy_(loss)
    stats[2:3].copy_(correct.sum(dtype=torch.float32).reshape(1))
    if world > 1:
        # flat = [prefix projection | tied matrix | positions | final norm | layer 0 .. layer L-1 | spare]: layers >= split are a contiguous tail
        views = bucket.views
        off_split = (views[4 + 6 * split].data_ptr() - flat.data_ptr()) // 4
        with torch.cuda.stream(side):
            side.wait_event(ev)
            early = dist.all_reduce(flat[off_split:bucket.total], op=dist.ReduceOp.SUM, group=group, async_op=True)
        dist.all_reduce(flat[:off_split], op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(flat[bucket.total:bucket.total + GradBucket.SPARE], op=dist.ReduceOp.SUM, group=group)
        early.wait()
        GradBucket.in_place_reductions += 1
    optimizer.max_grad_norm = float(gradient_clip)
    out = optimizer.step(flat_grads=flat, stats=stats)
    stats = stats.clone()
    return stats[0] / stats[1].clamp(min=1.0), stats[2], stats[1], out[0]


def train_step(model, optimizer, embed: torch.Tensor, target: torch.Tensor, mask: Optional[torch.Tensor], weight: Optional[torch.Tensor],
               noise=None, gradient_clip: float = 1.0, group: Optional[dist.ProcessGroup] = None):
    """One optimizer step on this rank's shard of the batch, mirroring train.py:1263-1286 with accum_factor folded into
    data parallelism: noise -> forward (loss_sum, loss_basis, correct) -> backward of loss_sum -> all-reduce of
    (gradients, loss_sum, loss_basis) -> normalise by the GLOBAL loss basis -> clip -> AdamW.step().
    Returns (global mean loss, global correct count, global token count, gradient norm).
    With a novic_b200.optim.FusedAdamW the step never synchronises with the host and overlaps the all-reduce with the backward pass;
    with a torch optimizer the reference's own sequence (autograd, clip_grad_norm_, optimizer.step) runs on the module's parameters."""
    from .optim import FusedAdamW
    if isinstance(optimizer, FusedAdamW):
        return _train_step_fused(model, optimizer, embed, target, mask, weight, noise, gradient_clip, group)
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if noise is not None:
        noise.stream_offset = dist.get_rank(group) if world > 1 else 0
        embed = noise(embed)
    model.dropout_seed_offset = (dist.get_rank(group) if world > 1 else 0) * 0x51ED2701
    optimizer.zero_grad(set_to_none=True)
    _, padding, loss_sum, loss_basis, correct = model(embed, target, mask, weight, True, True, False, None)
    loss_sum.backward()                                   # gradients of the local loss SUM (additive across shards)
    stats = torch.stack((loss_sum.detach().float(), loss_basis.detach().float(), correct.sum().float()))
    params = [p for p in model.parameters() if p.requires_grad]
    if world > 1:
        stats = allreduce_gradients(params, stats, group, bucket=getattr(model, "_grad_bucket", None))
    inv_basis = 1.0 / stats[1].clamp(min=1.0)
    torch._foreach_mul_([p.grad for p in params if p.grad is not None], inv_basis)
    norm = torch.nn.utils.clip_grad_norm_(params, max_norm=gradient_clip, error_if_nonfinite=True) if gradient_clip > 0 else None
    optimizer.step()
    return stats[0] * inv_basis, stats[2], stats[1], norm
