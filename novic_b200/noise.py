"""Embedding-noise modules with the reference's names, factory and in-place contract (embedding_noise.py:15-172),
implemented as ONE fused CUDA kernel (draw + perturb + renormalise) instead of 8-12 elementwise launches.

`EmbeddingNoise.create(scheme, embed_dim, vec_norm, angle_min, angle_max, angle_std, mix_ratio)` mirrors
embedding_noise.py:17-41.  The random draws come from a Philox stream keyed by (torch.initial_seed(), per-module
call counter): the distribution is the reference's, the stream is not torch's (bit parity with torch's RNG is
impossible; parity of the arithmetic is tested through `apply_predrawn`).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch

from . import _abi

_SCHEMES = {'gausselem': 0, 'gaussvec': 1, 'gaussangle': 2, 'uniformangle': 3, 'gausselemuniformangle': 4}


class EmbeddingNoise(torch.nn.Module):

    @staticmethod
    def create(scheme: str, embed_dim: int, vec_norm: float, angle_min: float, angle_max: float, angle_std: float,
               mix_ratio: float) -> Optional["EmbeddingNoise"]:
        if not scheme:
            return None
        key = scheme.lower()
        if key == 'gausselem':
            return GaussElemNoise(embed_dim=embed_dim, vec_norm=vec_norm)
        if key == 'gaussvec':
            return GaussVecNoise(embed_dim=embed_dim, vec_norm=vec_norm)
        if key == 'gaussangle':
            return GaussAngleNoise(embed_dim=embed_dim, angle_std=angle_std, angle_max=angle_max)
        if key == 'uniformangle':
            return UniformAngleNoise(embed_dim=embed_dim, angle_min=angle_min, angle_max=angle_max)
        if key == 'gausselemuniformangle':
            return GaussElemUniformAngleNoise(embed_dim=embed_dim, vec_norm=vec_norm, angle_min=angle_min, angle_max=angle_max, mix_ratio=mix_ratio)
        raise ValueError(f"Unsupported embedding noise type: {scheme}")

    def __init__(self, scheme: str, embed_dim: int, vec_norm: float = 0.0, angle_min: float = 0.0, angle_max: float = 0.0,
                 angle_std: float = 0.0, mix_ratio: float = 0.0):
        super().__init__()
        self.scheme = scheme
        self.embed_dim = embed_dim
        self._cfg = _abi.NovicNoiseCfg(scheme=_SCHEMES[scheme.lower()], embed_dim=embed_dim, vec_norm=vec_norm, angle_min=angle_min,
                                       angle_max=angle_max, angle_std=angle_std, mix_ratio=mix_ratio)
        self._calls = 0
        # Philox stream = (torch.initial_seed() + stream_offset, call counter).  Data-parallel training sets stream_offset to the rank
        # (novic_b200/dist.py) so that shards do not draw identical noise; a resumed run sets `call_counter` to its global step so that it
        # continues the sequence instead of replaying it (the module has no state-dict entries, like the reference's).
        self.stream_offset = 0

    @property
    def call_counter(self) -> int:
        return self._calls

    @call_counter.setter
    def call_counter(self, value: int) -> None:
        self._calls = int(value)

    def _check(self, embed: torch.Tensor) -> None:
        if embed.device.type != 'cuda':
            raise RuntimeError("novic_b200 embedding noise runs on CUDA only (no CPU path)")
        if embed.dtype != torch.float32 or embed.ndim != 2 or embed.shape[1] != self.embed_dim or not embed.is_contiguous():
            raise ValueError("embed must be a contiguous fp32 B x F tensor of unit vectors")

    def forward(self, embed: torch.Tensor) -> torch.Tensor:
        # embed = B x F unit vectors, modified in place and returned (embedding_noise.py:48-52)
        self._check(embed)
        self._calls += 1
        with torch.cuda.device(embed.device):
            _abi.check(_abi.lib().novic_noise_apply(C.byref(self._cfg), embed.data_ptr(), embed.shape[0],
                                                    (torch.initial_seed() + 0x9E3779B97F4A7C15 * int(self.stream_offset)) & 0xFFFFFFFFFFFFFFFF, self._calls,
                                                    torch.cuda.current_stream(embed.device).cuda_stream))
        return embed

    def apply_predrawn(self, embed: torch.Tensor, normals_a: torch.Tensor, normals_b: Optional[torch.Tensor] = None,
                       row_a: Optional[torch.Tensor] = None, row_b: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Deterministic variant: the caller supplies the N(0,1) / U[0,1) draws (see include/novic_b200.h)."""
        self._check(embed)
        for t in (normals_a, normals_b, row_a, row_b):
            if t is not None and (t.device != embed.device or t.dtype != torch.float32 or not t.is_contiguous()):
                raise ValueError("pre-drawn tensors must be contiguous fp32 on the embedding's device")
        p = lambda t: None if t is None else t.data_ptr()
        with torch.cuda.device(embed.device):
            _abi.check(_abi.lib().novic_noise_apply_predrawn(C.byref(self._cfg), embed.data_ptr(), embed.shape[0], p(normals_a), p(normals_b),
                                                             p(row_a), p(row_b), torch.cuda.current_stream(embed.device).cuda_stream))
        return embed


class GaussElemNoise(EmbeddingNoise):
    def __init__(self, embed_dim: int, vec_norm: float):
        self.vec_norm = vec_norm
        self.elem_std = vec_norm / math.sqrt(embed_dim)
        if self.elem_std <= 0:
            raise ValueError(f"Element noise standard deviation must be positive: {self.elem_std:.3g}")
        super().__init__('GaussElem', embed_dim, vec_norm=vec_norm)

    def extra_repr(self) -> str:
        return f"embed_dim={self.embed_dim}, vec_norm={self.vec_norm:.3g}, elem_std={self.elem_std:.3g}"


class GaussVecNoise(EmbeddingNoise):
    def __init__(self, embed_dim: int, vec_norm: float):
        self.vec_norm = vec_norm
        if vec_norm <= 0:
            raise ValueError(f"Vector noise norm must be positive: {vec_norm:.3g}")
        super().__init__('GaussVec', embed_dim, vec_norm=vec_norm)

    def extra_repr(self) -> str:
        return f"embed_dim={self.embed_dim}, vec_norm={self.vec_norm:.3g}"


class AngleNoise(EmbeddingNoise):
    pass


class GaussAngleNoise(AngleNoise):
    def __init__(self, embed_dim: int, angle_std: float, angle_max: float):
        self.angle_std, self.angle_max = angle_std, angle_max
        if math.radians(angle_std) <= 0 or math.radians(angle_max) <= 0:
            raise ValueError("Angular noise standard deviation and maximum value must both be positive")
        super().__init__('GaussAngle', embed_dim, angle_std=angle_std, angle_max=angle_max)

    def extra_repr(self) -> str:
        return f"embed_dim={self.embed_dim}, angle_std={self.angle_std:.3g}\xB0, angle_max={self.angle_max:.3g}\xB0"


class UniformAngleNoise(AngleNoise):
    def __init__(self, embed_dim: int, angle_min: float, angle_max: float):
        self.angle_min, self.angle_max = angle_min, angle_max
        if angle_min > angle_max:
            raise ValueError("Minimum angular noise must be smaller than maximum angular noise")
        super().__init__('UniformAngle', embed_dim, angle_min=angle_min, angle_max=angle_max)

    def extra_repr(self) -> str:
        return f"embed_dim={self.embed_dim}, angle_min={self.angle_min:.3g}\xB0, angle_max={self.angle_max:.3g}\xB0"


class GaussElemUniformAngleNoise(EmbeddingNoise):
    def __init__(self, embed_dim: int, vec_norm: float, angle_min: float, angle_max: float, mix_ratio: float):
        self.vec_norm, self.angle_min, self.angle_max, self.mix_ratio = vec_norm, angle_min, angle_max, mix_ratio
        if vec_norm / math.sqrt(embed_dim) <= 0:
            raise ValueError("Element noise standard deviation must be positive")
        if angle_min > angle_max:
            raise ValueError("Minimum angular noise must be smaller than maximum angular noise")
        if mix_ratio < 0 or mix_ratio > 1:
            raise ValueError(f"Mix ratio must be in the range [0, 1]: {mix_ratio:.3g}")
        super().__init__('GaussElemUniformAngle', embed_dim, vec_norm=vec_norm, angle_min=angle_min, angle_max=angle_max, mix_ratio=mix_ratio)

    def extra_repr(self) -> str:
        return f"mix_ratio={self.mix_ratio:.3g}"
