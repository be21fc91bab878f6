"""Image encoder in front of the decoder (BASELINE config #5: CLIP ViT-H/14-378 + decoder end to end).

The reference gets its image embeddings from `open_clip`'s `model.encode_image(images, normalize=False)` and normalises them in fp32
(embedders.py:752-764).  `ImageEncoder` offers the same call on the same parameter names (`visual.*` of open_clip's
`VisionTransformer`), computed by libnovic_b200.so (novic_vit_encode: tcgen05 GEMMs with fused bias / QuickGELU / residual epilogues,
tensor-core flash attention, LayerNorm kernels); `EncoderDecoder` hands the embeddings to `PrefixedIterDecoder.generate` without leaving
the device.  open_clip_torch is an un-vendored dependency of the reference, so the architecture is an assumption spelt out in
DESIGN.md and parity is pinned only on the independent restatement oracle/vit_oracle.py ("parity unpinned").  No CPU path.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import math

import numpy as np
import torch
import torch.nn as nn

from . import _abi


@dataclasses.dataclass(frozen=True)
class VitDims:
    """open_clip `ViT-H-14-378-quickgelu` (DFN5B-CLIP-ViT-H-14-378, config/train.yaml:104) as assumed in DESIGN.md."""
    image_size: int = 378
    patch_size: int = 14
    width: int = 1280
    layers: int = 32
    heads: int = 16
    mlp_dim: int = 5120
    out_dim: int = 1024
    ln_eps: float = 1e-5

    @property
    def tokens(self) -> int:
        return (self.image_size // self.patch_size) ** 2 + 1


# ---- parameter holders: they exist so that state_dict() has open_clip's key names; none has a forward() ----
class _Mlp(nn.Module):
    def __init__(self, width, mlp_dim):
        super().__init__()
        self.c_fc = nn.Linear(width, mlp_dim)
        self.c_proj = nn.Linear(mlp_dim, width)


class _Block(nn.Module):
    def __init__(self, width, heads, mlp_dim, eps):
        super().__init__()
        self.ln_1 = nn.LayerNorm(width, eps=eps)
        self.attn = nn.MultiheadAttention(width, heads)      # holder of in_proj_weight / in_proj_bias / out_proj.{weight,bias}
        self.ln_2 = nn.LayerNorm(width, eps=eps)
        self.mlp = _Mlp(width, mlp_dim)


class _Blocks(nn.Module):
    def __init__(self, d: VitDims):
        super().__init__()
        self.resblocks = nn.ModuleList([_Block(d.width, d.heads, d.mlp_dim, d.ln_eps) for _ in range(d.layers)])


class _Visual(nn.Module):
    def __init__(self, d: VitDims):
        super().__init__()
        self.conv1 = nn.Conv2d(3, d.width, kernel_size=d.patch_size, stride=d.patch_size, bias=False)
        self.class_embedding = nn.Parameter(torch.empty(d.width))
        self.positional_embedding = nn.Parameter(torch.empty(d.tokens, d.width))
        self.ln_pre = nn.LayerNorm(d.width, eps=d.ln_eps)
        self.transformer = _Blocks(d)
        self.ln_post = nn.LayerNorm(d.width, eps=d.ln_eps)
        self.proj = nn.Parameter(torch.empty(d.width, d.out_dim))


class ImageEncoder(nn.Module):
    def __init__(self, dims: VitDims = VitDims(), images_per_chunk: int = 64):
        super().__init__()
        if dims.width != dims.heads * 80 or dims.width not in (640, 1280):
            raise ValueError("ImageEncoder (novic_b200) supports heads of 80 channels and width 640 or 1280")
        if dims.layers > _abi.NOVIC_VIT_MAX_LAYERS:
            raise ValueError(f"at most {_abi.NOVIC_VIT_MAX_LAYERS} blocks")
        self.dims = dims
        self.images_per_chunk = int(images_per_chunk)
        self.visual = _Visual(dims)
        self.reset_parameters()
        self._handles: dict[int, dict] = {}

    @torch.no_grad()
    def reset_parameters(self) -> None:
        """Random initialisation in the style of CLIP's (scale = width^-0.5 for embeddings and projection, attention / MLP weights
        scaled by width and depth); the benchmark only needs well-conditioned random weights of the right shapes."""
        d = self.dims
        scale = d.width ** -0.5
        v = self.visual
        nn.init.normal_(v.conv1.weight, std=(3 * d.patch_size ** 2) ** -0.5)
        nn.init.normal_(v.class_embedding, std=scale)
        nn.init.normal_(v.positional_embedding, std=scale)
        nn.init.normal_(v.proj, std=scale)
        proj_std = scale * (2 * d.layers) ** -0.5
        for b in v.transformer.resblocks:
            nn.init.normal_(b.attn.in_proj_weight, std=scale)
            nn.init.normal_(b.attn.out_proj.weight, std=proj_std)
            nn.init.normal_(b.mlp.c_fc.weight, std=(2 * d.width) ** -0.5)
            nn.init.normal_(b.mlp.c_proj.weight, std=proj_std)
            for t in (b.attn.in_proj_bias, b.attn.out_proj.bias, b.mlp.c_fc.bias, b.mlp.c_proj.bias):
                nn.init.zeros_(t)

    def _weight_tensors(self) -> list[torch.Tensor]:
        v = self.visual
        ts = [v.conv1.weight, v.class_embedding, v.positional_embedding, v.ln_pre.weight, v.ln_pre.bias, v.ln_post.weight, v.ln_post.bias, v.proj]
        for b in v.transformer.resblocks:
            ts += [b.ln_1.weight, b.ln_1.bias, b.attn.in_proj_weight, b.attn.in_proj_bias, b.attn.out_proj.weight, b.attn.out_proj.bias,
                   b.ln_2.weight, b.ln_2.bias, b.mlp.c_fc.weight, b.mlp.c_fc.bias, b.mlp.c_proj.weight, b.mlp.c_proj.bias]
        return ts

    def _state(self, device: torch.device) -> dict:
        if device.type != "cuda":
            raise RuntimeError(f"novic_b200.ImageEncoder computes on CUDA (sm_100a) only; there is no CPU path. Got tensors on {device}.")
        lib = _abi.lib()
        idx = device.index if device.index is not None else torch.cuda.current_device()
        st = self._handles.get(idx)
        d = self.dims
        with torch.cuda.device(idx):
            if st is None:
                cfg = _abi.NovicVitCfg(d.image_size, d.patch_size, d.width, d.layers, d.heads, d.mlp_dim, d.out_dim, d.ln_eps)
                handle = C.c_void_p()
                _abi.check(lib.novic_vit_create(C.byref(cfg), C.byref(handle)))
                st = dict(handle=handle, wbuf=None, wkey=None, ws=None)
                self._handles[idx] = st
            tensors = self._weight_tensors()
            wkey = tuple((t.data_ptr(), 0 if t.is_inference() else t._version) for t in tensors)
            if st["wkey"] != wkey:
                for t in tensors:
                    if t.device.type != "cuda" or t.device.index != idx or t.dtype != torch.float32 or not t.is_contiguous():
                        raise RuntimeError(f"encoder parameters must be contiguous fp32 tensors on the input's CUDA device (found {t.dtype} on {t.device})")
                if st["wbuf"] is None:
                    st["wbuf"] = torch.empty(lib.novic_vit_weight_bytes(st["handle"]), dtype=torch.uint8, device=device)
                w = _abi.NovicVitWeights()
                (w.conv1, w.class_embedding, w.positional_embedding, w.ln_pre_w, w.ln_pre_b, w.ln_post_w, w.ln_post_b, w.proj) = (t.data_ptr() for t in tensors[:8])
                names = ("ln1_w", "ln1_b", "in_proj_w", "in_proj_b", "out_proj_w", "out_proj_b", "ln2_w", "ln2_b", "fc_w", "fc_b", "cproj_w", "cproj_b")
                for i in range(d.layers):
                    for j, name in enumerate(names):
                        getattr(w, name)[i] = tensors[8 + 12 * i + j].data_ptr()
                _abi.check(lib.novic_vit_set_weights(st["handle"], C.byref(w), st["wbuf"].data_ptr(), st["wbuf"].numel(), torch.cuda.current_stream(idx).cuda_stream))
                st["wkey"] = wkey
        return st

    def encode_image(self, images: torch.Tensor, normalize: bool = False) -> torch.Tensor:
        """images [B, 3, S, S] fp32 (preprocessed) on the CUDA device -> [B, out_dim] fp32 (open_clip's encode_image)."""
        d = self.dims
        assert images.ndim == 4 and images.shape[1] == 3 and images.shape[2] == images.shape[3] == d.image_size and images.dtype == torch.float32
        images = images.contiguous()
        dev = images.device
        st = self._state(dev)
        lib = _abi.lib()
        B = images.shape[0]
        chunk = max(1, min(self.images_per_chunk, B))
        need = lib.novic_vit_workspace_bytes(st["handle"], chunk)
        if st["ws"] is None or st["ws"].numel() < need:
            st["ws"] = None
            st["ws"] = torch.empty(need, dtype=torch.uint8, device=dev)
        out = torch.empty((B, d.out_dim), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _abi.check(lib.novic_vit_encode(st["handle"], images.data_ptr(), B, out.data_ptr(), int(bool(normalize)), chunk, st["ws"].data_ptr(),
                                            st["ws"].numel(), torch.cuda.current_stream(dev).cuda_stream))
        return out

    forward = encode_image

    def __del__(self):
        try:
            lib = _abi.lib()
            for st in getattr(self, "_handles", {}).values():
                lib.novic_vit_destroy(st["handle"])
        except Exception:
            pass


class EncoderDecoder(nn.Module):
    """Images -> labels on the device: Embedder.inference_image (embedders.py:759-764: encode_image(normalize=False) + fp32 normalise)
    followed by GenerationTask.generate's greedy call (infer.py:567-576)."""

    def __init__(self, encoder: ImageEncoder, decoder):
        super().__init__()
        self.encoder, self.decoder = encoder, decoder

    def embed(self, images: torch.Tensor) -> torch.Tensor:
        return self.encoder.encode_image(images, normalize=True)

    def generate(self, images: torch.Tensor, temperature: float = 1.0, length_alpha: float = 0.0, guide_targets=None, guide_renorm: bool = False):
        target, padding, _, _, _, score = self.decoder.generate(self.embed(images), False, True, temperature, length_alpha, None, guide_targets, guide_renorm)
        return target.unsqueeze(1), padding.unsqueeze(1), score.unsqueeze(1)


def synth_vit_state_dict(dims: VitDims = VitDims(), seed: int = 7) -> dict:
    """Random `visual.*` state dict, reproducible from (seed, shape) alone (numpy PCG64): CLIP-style scales, LayerNorm gains jittered
    around 1 and small non-zero biases everywhere so that a kernel that forgot one of them is caught."""
    rng = np.random.default_rng(seed)
    W, L, M, O, T, P = dims.width, dims.layers, dims.mlp_dim, dims.out_dim, dims.tokens, dims.patch_size
    scale = W ** -0.5
    proj_std = scale * (2 * L) ** -0.5

    def n(shape, std):
        return torch.from_numpy((rng.standard_normal(shape) * std).astype(np.float32))

    def gain(k):
        return torch.from_numpy((1.0 + 0.2 * rng.standard_normal(k)).astype(np.float32))
    sd = {"visual.conv1.weight": n((W, 3, P, P), (3 * P * P) ** -0.5), "visual.class_embedding": n((W,), scale),
          "visual.positional_embedding": n((T, W), scale), "visual.ln_pre.weight": gain(W), "visual.ln_pre.bias": n((W,), 0.05),
          "visual.ln_post.weight": gain(W), "visual.ln_post.bias": n((W,), 0.05), "visual.proj": n((W, O), scale)}
    for i in range(L):
        p = f"visual.transformer.resblocks.{i}."
        sd[p + "ln_1.weight"] = gain(W); sd[p + "ln_1.bias"] = n((W,), 0.05)
        sd[p + "attn.in_proj_weight"] = n((3 * W, W), scale * 2.0); sd[p + "attn.in_proj_bias"] = n((3 * W,), 0.05)
        sd[p + "attn.out_proj.weight"] = n((W, W), proj_std * 4.0); sd[p + "attn.out_proj.bias"] = n((W,), 0.02)
        sd[p + "ln_2.weight"] = gain(W); sd[p + "ln_2.bias"] = n((W,), 0.05)
        sd[p + "mlp.c_fc.weight"] = n((M, W), (2 * W) ** -0.5 * 2.0); sd[p + "mlp.c_fc.bias"] = n((M,), 0.05)
        sd[p + "mlp.c_proj.weight"] = n((W, M), proj_std * 4.0); sd[p + "mlp.c_proj.bias"] = n((W,), 0.02)
    return sd


def synth_images(batch: int, dims: VitDims = VitDims(), seed: int = 11) -> torch.Tensor:
    """Synthetic preprocessed images (zero mean, unit variance per channel, with some spatial structure)."""
    rng = np.random.default_rng(seed)
    S = dims.image_size
    base = rng.standard_normal((batch, 3, S // 7 + 1, S // 7 + 1)).astype(np.float32)
    img = np.repeat(np.repeat(base, 7, axis=2), 7, axis=3)[:, :, :S, :S]
    img = 0.7 * img + 0.7 * rng.standard_normal((batch, 3, S, S)).astype(np.float32)
    return torch.from_numpy(np.ascontiguousarray(img))
