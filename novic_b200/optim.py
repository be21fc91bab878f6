"""Fused optimizer tail of the training step: global-norm gradient clipping + AdamW over flat buffers, on the device.

The reference's step is `clip_grad_norm_(params, max_norm, error_if_nonfinite=True)` followed by `torch.optim.AdamW(fused=True).step()`
(train.py:1281-1286) on the parameter groups of train.py:1108-1119 (weight decay on tensors with >= 2 dimensions only).  The clip reads
the norm back to the host, so the host can never run ahead of the GPU.  `FusedAdamW` keeps parameters, both moments and (through
novic_b200.training) the gradients as ONE flat fp32 buffer each and runs the whole tail as three kernels of libnovic_b200.so
(`novic_adamw_step`: sum of squares -> clip coefficient -> update) without any host synchronisation; the 1 / loss_basis normalisation
of the gradients (train.py:1272) is folded into the same pass.

It is a `torch.optim.Optimizer`: `param_groups` (with 'lr', 'betas', 'eps', 'weight_decay' - LR schedulers work on it), `state` with the
keys torch's AdamW uses ('step', 'exp_avg', 'exp_avg_sq', the moments being views of the flat buffers), `state_dict()` /
`load_state_dict()`, `zero_grad()`.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _abi

CHUNK = 512   # granularity of the weight-decay flags (csrc/optim.cuh kOptChunk)


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, model, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 weight_decay_1d: bool = False, max_grad_norm: float = 0.0):
        """model: a novic_b200.PrefixedIterDecoder already on its CUDA device.  Its parameters are re-seated as views of one flat buffer
        (same objects, same values, same state-dict keys).  weight_decay_1d / the two groups follow train.py:1108-1119."""
        params = model._weight_tensors()
        if any(p.device.type != "cuda" for p in params):
            raise RuntimeError("FusedAdamW needs the model on a CUDA device (call .to(device) first): novic_b200 has no CPU path")
        if any(p.dtype != torch.float32 for p in params):
            raise ValueError("FusedAdamW supports fp32 parameters")
        if {id(p) for p in params} != {id(p) for p in model.parameters()}:
            raise ValueError("the model has parameters outside its weight list")
        dev = params[0].device
        sizes = [p.numel() for p in params]
        if any(n % CHUNK for n in sizes):
            raise ValueError(f"every parameter tensor must be a multiple of {CHUNK} elements")
        total = sum(sizes)
        flat = torch.empty(total, dtype=torch.float32, device=dev)
        flags = torch.zeros(total // CHUNK, dtype=torch.uint8)
        off = 0
        with torch.no_grad():
            for p, n in zip(params, sizes):
                flat[off:off + n].copy_(p.detach().reshape(-1))
                p.data = flat[off:off + n].view(p.shape)          # the Parameter object stays, its storage becomes a slice of `flat`
                if weight_decay_1d or p.dim() >= 2:
                    flags[off // CHUNK:(off + n) // CHUNK] = 1
                off += n
        one_d = [p for p in params if p.dim() < 2]
        n_d = [p for p in params if p.dim() >= 2]
        if weight_decay_1d:
            groups = [{"params": params, "weight_decay": weight_decay}]
        else:
            groups = [{"params": one_d, "weight_decay": 0.0}, {"params": n_d, "weight_decay": weight_decay}]
        super().__init__(groups, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self.model = model
        self.max_grad_norm = float(max_grad_norm)
        self._params, self._sizes, self._total = params, sizes, total
        self.flat_params = flat
        self.flat_exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self._decay_flags = flags.to(dev)
        self._scratch = torch.empty(_abi.lib().novic_adamw_scratch_bytes(), dtype=torch.uint8, device=dev)
        self.last = torch.zeros(4, dtype=torch.float32, device=dev)       # [grad norm, clip coefficient, gradient factor, non-finite flag] of the last step
        self._step = 0
        self._own_grads: Optional[torch.Tensor] = None
        off = 0
        for p, n in zip(params, sizes):
            self.state[p] = {"step": torch.tensor(0.0), "exp_avg": self.flat_exp_avg[off:off + n].view(p.shape),
                             "exp_avg_sq": self.flat_exp_avg_sq[off:off + n].view(p.shape)}
            off += n
        model.refresh_weights()

    # ------------------------------------------------------------------------------------------------------------------
    def _weight_decay(self) -> float:
        return max(float(g["weight_decay"]) for g in self.param_groups)

    def _flat_grads(self) -> torch.Tensor:
        """The gradients as one flat buffer in parameter order: the training step's own bucket when `.grad` are its views (no copy),
        else a packed copy."""
        bucket = getattr(self.model, "_grad_bucket", None)
        grads = [p.grad for p in self._params]
        if any(g is None for g in grads):
            raise RuntimeError("FusedAdamW.step(): a parameter has no gradient")
        if bucket is not None and bucket.covers(grads) and bucket.in_parameter_order(grads):
            return bucket.flat
        if self._own_grads is None:
            self._own_grads = torch.empty(self._total, dtype=torch.float32, device=self.flat_params.device)
        torch.cat([g.reshape(-1) for g in grads], out=self._own_grads)
        return self._own_grads

    @torch.no_grad()
    def step(self, closure=None, *, flat_grads: Optional[torch.Tensor] = None, stats: Optional[torch.Tensor] = None):
        """One update.  flat_grads: the gradient bucket (d loss_sum / d theta in parameter order; default: taken from `.grad`);
        stats: device tensor [loss_sum, loss_basis, ...] whose loss_basis normalises the gradients on the device (default: the
        gradients are used as they are, e.g. after `(loss_sum / loss_basis).backward()`).  Returns self.last (device tensor, no
        synchronisation): gradient norm, clip coefficient, gradient factor, non-finite flag."""
        if closure is not None:
            raise NotImplementedError("FusedAdamW does not re-evaluate closures")
        g = self._flat_grads() if flat_grads is None else flat_grads
        if g.dtype != torch.float32 or g.numel() < self._total or not g.is_contiguous() or g.device != self.flat_params.device:
            raise ValueError("flat_grads must be a contiguous fp32 CUDA buffer of at least the parameter count")
        if stats is not None and (stats.dtype != torch.float32 or stats.numel() < 2 or stats.device != g.device):
            raise ValueError("stats must be an fp32 tensor [loss_sum, loss_basis, ...] on the gradients' device")
        self._step += 1
        grp = self.param_groups[0]
        beta1, beta2 = grp["betas"]
        cfg = _abi.NovicAdamW(float(grp["lr"]), float(beta1), float(beta2), float(grp["eps"]), self._weight_decay(), self.max_grad_norm, self._step)
        dev = self.flat_params.device
        with torch.cuda.device(dev):
            _abi.check(_abi.lib().novic_adamw_step(
                C.byref(cfg), self.flat_params.data_ptr(), g.data_ptr(), self.flat_exp_avg.data_ptr(), self.flat_exp_avg_sq.data_ptr(), self._total,
                self._decay_flags.data_ptr(), None if stats is None else stats.data_ptr(), self._scratch.data_ptr(), self._scratch.numel(),
                self.last.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        for p in self._params:
            self.state[p]["step"] = torch.tensor(float(self._step))
        self.model.refresh_weights()          # the bf16 operand copies are re-packed on the next call
        return self.last

    def grad_norm(self) -> float:
        """Gradient norm of the last step (host synchronisation); raises like clip_grad_norm_(error_if_nonfinite=True) if it was not
        finite - the update of that step was skipped on the device."""
        norm, _, _, bad = self.last.tolist()
        if bad:
            raise RuntimeError("The total norm of the gradients of the last step was non-finite, so it was not applied")
        return norm

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        # torch replaced the state tensors by copies: move the values back into the flat buffers and re-seat the views
        off, step = 0, 0
        for p, n in zip(self._params, self._sizes):
            st = self.state[p]
            self.flat_exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
            self.flat_exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
            st["exp_avg"] = self.flat_exp_avg[off:off + n].view(p.shape)
            st["exp_avg_sq"] = self.flat_exp_avg_sq[off:off + n].view(p.shape)
            step = max(step, int(float(st["step"])))
            off += n
        self._step = step
