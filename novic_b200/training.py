"""Training-mode forward/backward of the decoder (train.py:1263-1286).

`PrefixedIterDecoder.forward` in training mode with autograd enabled routes here: one library call
(`novic_train_fwd_bwd`) runs the teacher-forced forward *and* the full backward and leaves d(loss_sum)/d(parameter)
for all parameters in fp32 buffers; the autograd Function hands them to torch scaled by the incoming gradient of
loss_sum, so the reference's own `loss.backward()`, `clip_grad_norm_` and `torch.optim.AdamW.step()`
(train.py:1273-1286) work unchanged on the module's parameters.  Dropout (input + layer) is applied inside the library call with
hash-generated masks (novic_set_dropout); the seed of each call is drawn from torch's CPU generator.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _abi


def _grad_struct(model, grads: list[torch.Tensor]) -> _abi.NovicWeights:
    g = _abi.NovicWeights()
    g.embed_mlp, g.tok_embed, g.pos_embed, g.final_norm = (t.data_ptr() for t in grads[:4])
    for i in range(len(model.transformer.layers)):
        base = 4 + 6 * i
        g.in_proj[i], g.out_proj[i], g.linear1[i], g.linear2[i], g.norm1[i], g.norm2[i] = (t.data_ptr() for t in grads[base:base + 6])
    return g


def grad_bucket(model, params):
    """The flat fp32 buffer the gradients of a training step are views of, persistent across steps (stable addresses keep the library's
    captured graph of the step valid).  A fresh buffer is used while `.grad` of a previous step still lives in the persistent one
    (gradient accumulation without zero_grad: autograd would otherwise add a buffer onto itself) or while an earlier forward's
    gradients have not been handed to autograd yet (two forwards before one backward)."""
    from .dist import GradBucket
    dev = params[0].device
    sizes = [p.numel() for p in params]
    total = sum(sizes)
    cached = getattr(model, "_grad_bucket_persistent", None)
    usable = cached is not None and cached.flat.device == dev and cached.total == total
    if usable:
        lo, hi = cached.flat.data_ptr(), cached.flat.data_ptr() + cached.flat.numel() * 4
        if not getattr(cached, "busy", False) and not any(p.grad is not None and lo <= p.grad.data_ptr() < hi for p in params):
            return cached
    flat = torch.empty(total + GradBucket.SPARE, dtype=torch.float32, device=dev)
    grads, off = [], 0
    for p, n in zip(params, sizes):
        grads.append(flat[off:off + n].view(p.shape))
        off += n
    bucket = GradBucket(flat, grads)
    bucket.views = grads
    if not usable:
        model._grad_bucket_persistent = bucket
    return bucket


def fwd_bwd(model, embed, target, padding, weight, M, split_layer: int = -1, split_event=None):
    """One library call: teacher-forced forward + loss + full backward (novic_train_fwd_bwd_ex).  Returns (loss [2] = {loss_sum,
    loss_basis}, correct u8 [A, C], effective padding u8 [A, C], GradBucket holding d(loss_sum) / d(parameter) for all parameters).
    split_layer > 0 with a torch.cuda.Event: the event is recorded once the gradients of layers >= split_layer are final."""
    params = model._weight_tensors()
    dev = embed.device
    st = model._state(dev, refresh=True)   # parameters may have been stepped by a fused optimizer since the last call
    lib = _abi.lib()
    B = embed.shape[0]
    A, Ct = target.shape
    need = lib.novic_train_workspace_bytes(st['handle'], B, M, Ct)
    ws = st.get('train_ws')
    if ws is None or ws.numel() < need:
        st['train_ws'] = None
        st['train_ws'] = ws = torch.empty(need, dtype=torch.uint8, device=dev)
    bucket = grad_bucket(model, params)
    grads = bucket.views
    model._grad_bucket = bucket
    V = model.target_config.vocab_size
    if grads[1].shape[0] > V:
        grads[1][V:].zero_()    # vocab_quant rows never receive a gradient (the library writes the V used rows only)
    bucket.flat[bucket.total:].zero_()
    loss = torch.empty(2, dtype=torch.float32, device=dev)
    correct = torch.empty((A, Ct), dtype=torch.uint8, device=dev)
    pad_out = torch.empty((A, Ct), dtype=torch.uint8, device=dev)
    gs = _grad_struct(model, grads)
    p_in, p_layer = dropout_probs(model)
    seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if (p_in > 0 or p_layer > 0) else 0
    seed += int(getattr(model, "dropout_seed_offset", 0))     # data parallelism: a different mask stream per rank (novic_b200/dist.py)
    _abi.check(lib.novic_set_dropout(st['handle'], float(p_in), float(p_layer), seed))
    model._last_dropout = (p_in, p_layer, seed)      # tests replay the masks from this
    ev_handle = None
    if split_layer > 0 and split_event is not None:
        ev_handle = split_event.cuda_event
    with torch.cuda.device(dev):
        _abi.check(lib.novic_train_fwd_bwd_ex(
            st['handle'], embed.data_ptr(), B, M, target.data_ptr(), None if padding is None else padding.data_ptr(),
            None if weight is None else weight.data_ptr(), Ct, loss.data_ptr(), correct.data_ptr(), pad_out.data_ptr(), C.byref(gs),
            ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream, split_layer if ev_handle is not None else -1, ev_handle))
    return loss, correct, pad_out, bucket


class TrainStep(torch.autograd.Function):
    """Inputs: (model, embed, target [A, C], padding u8 [A, C] | None, weight [A] | None, M, *parameters).
    Outputs: loss_sum (0-dim fp32), loss_basis (0-dim fp32), correct (u8 [A, C]), effective padding (u8 [A, C])."""

    @staticmethod
    def forward(ctx, model, embed, target, padding, weight, M, *params):
        # the gradients are views of one flat buffer: autograd passes the views on to `.grad` as they are, so the data-parallel
        # all-reduce (dist.allreduce_gradients) and the fused optimizer (optim.FusedAdamW) run on the buffer in place
        loss, correct, pad_out, bucket = fwd_bwd(model, embed, target, padding, weight, M)
        bucket.busy = True
        ctx.bucket = bucket
        ctx.grads = bucket.views
        ctx.mark_non_differentiable(correct, pad_out)
        return loss[0], loss[1], correct, pad_out

    @staticmethod
    def backward(ctx, g_loss_sum, g_loss_basis, g_correct, g_pad):
        grads = ctx.grads
        ctx.grads = None
        ctx.bucket.busy = False
        if g_loss_sum is None:
            return (None,) * 6 + tuple(None for _ in grads)
        torch._foreach_mul_(grads, g_loss_sum)   # d(loss)/d(theta) = d(loss)/d(loss_sum) * d(loss_sum)/d(theta)
        return (None,) * 6 + tuple(grads)


def dropout_probs(model) -> tuple[float, float]:
    """(input, layer) dropout probabilities of a training step, from the nn.Dropout holders (utils.rescale_dropout mutates those)."""
    if not model.training:
        return 0.0, 0.0
    layers = model.transformer.layers
    return float(model.pos_embedding.dropout.p), float(layers[0].dropout.p) if len(layers) else 0.0


def train_forward(model, embed: torch.Tensor, target: torch.Tensor, padding: Optional[torch.Tensor], weight: Optional[torch.Tensor], M: int):
    params = model._weight_tensors()
    pad_u8 = None if padding is None else padding.contiguous().view(torch.uint8)
    w = None if weight is None else weight.contiguous()
    return TrainStep.apply(model, embed, target.contiguous(), pad_u8, w, M, *params)
