"""CPU restatement of the image encoder of BASELINE config #5 - PARITY UNPINNED.

The reference obtains image embeddings from `open_clip.create_model_and_transforms(...)` and calls
`model.encode_image(images, normalize=False)` followed by an fp32 normalise (embedders.py:752-764).  open_clip_torch==2.23
(requirements.txt:8) is neither vendored under /root/reference nor installed, and no weights or tests exist at that boundary, so
this file restates the PUBLISHED architecture of open_clip's `VisionTransformer` for the `ViT-H-14-378-quickgelu` configuration
from memory (an assumption the repository cannot verify offline):

    x = conv1(image)                      # 3 -> width, kernel = stride = patch, no bias           [B, width, np, np]
    x = cat(class_embedding, x.flatten) + positional_embedding                                     [B, np*np + 1, width]
    x = ln_pre(x)
    for block in resblocks:               # pre-LN residual blocks
        x = x + attn(ln_1(x))             # nn.MultiheadAttention(width, heads) with biases, no mask
        x = x + c_proj(quick_gelu(c_fc(ln_2(x))))      # quick_gelu(v) = v * sigmoid(1.702 v)
    pooled = ln_post(x[:, 0])             # class token
    return pooled @ proj                  # [width, out_dim], no bias

with width 1280, 32 blocks, 16 heads of 80, mlp 5120, patch 14 on 378 x 378 images (730 tokens), out_dim 1024, LayerNorm eps 1e-5.
State-dict keys follow open_clip (`visual.*`).  TEST INFRASTRUCTURE ONLY: nothing under novic_b200/ imports this module.
"""
from __future__ import annotations

import dataclasses
import math

import torch
import torch.nn.functional as F


@dataclasses.dataclass(frozen=True)
class VitCfg:
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


def encode_image(cfg: VitCfg, sd: dict, images: torch.Tensor, normalize: bool = False, bf16_operands: bool = False) -> torch.Tensor:
    """images [B, 3, S, S] fp32 -> [B, out_dim] fp32.
    bf16_operands: round every matrix-product operand (weights, LayerNorm outputs, q / k / v, attention probabilities and outputs, the MLP
    hidden rows) to bf16 first, as a kernel with bf16 tensor-core operands and fp32 accumulation sees them: a CUDA path must then agree to
    accumulation-order level, which separates a wrong kernel from honest rounding."""
    r = (lambda t: t.bfloat16().float()) if bf16_operands else (lambda t: t)
    B = images.shape[0]
    W, H = cfg.width, cfg.heads
    x = F.conv2d(r(images), r(sd["visual.conv1.weight"]), stride=cfg.patch_size)                 # B x W x np x np
    x = x.reshape(B, W, -1).permute(0, 2, 1)                                                      # B x np^2 x W
    cls = sd["visual.class_embedding"].to(x.dtype).expand(B, 1, W)
    x = torch.cat((cls, x), dim=1) + sd["visual.positional_embedding"]
    x = F.layer_norm(x, (W,), sd["visual.ln_pre.weight"], sd["visual.ln_pre.bias"], cfg.ln_eps)
    T = x.shape[1]
    for i in range(cfg.layers):
        p = f"visual.transformer.resblocks.{i}."
        y = r(F.layer_norm(x, (W,), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], cfg.ln_eps))
        qkv = r(F.linear(y, r(sd[p + "attn.in_proj_weight"]), sd[p + "attn.in_proj_bias"]))
        q, k, v = (t.reshape(B, T, H, W // H).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
        sc = q @ k.transpose(-1, -2) / math.sqrt(W // H)
        if bf16_operands:   # the kernel rounds the un-normalised exponentials to bf16 and divides by their fp32 sum afterwards
            e = torch.exp(sc - sc.amax(dim=-1, keepdim=True))
            att = (r(e) @ v) / e.sum(dim=-1, keepdim=True)
        else:
            att = torch.softmax(sc, dim=-1) @ v
        att = r(att.transpose(1, 2).reshape(B, T, W))
        x = x + F.linear(att, r(sd[p + "attn.out_proj.weight"]), sd[p + "attn.out_proj.bias"])
        y = r(F.layer_norm(x, (W,), sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], cfg.ln_eps))
        h = F.linear(y, r(sd[p + "mlp.c_fc.weight"]), sd[p + "mlp.c_fc.bias"])
        h = r(h * torch.sigmoid(1.702 * h))
        x = x + F.linear(h, r(sd[p + "mlp.c_proj.weight"]), sd[p + "mlp.c_proj.bias"])
    pooled = r(F.layer_norm(x[:, 0], (W,), sd["visual.ln_post.weight"], sd["visual.ln_post.bias"], cfg.ln_eps))
    out = pooled @ r(sd["visual.proj"])
    return F.normalize(out, dim=-1) if normalize else out
