"""Image encoder (BASELINE config #5) against the independent CPU restatement oracle/vit_oracle.py.

PARITY UNPINNED: open_clip_torch is an un-vendored dependency of the reference (requirements.txt:8), so the oracle restates the
published VisionTransformer from memory and cannot be checked against the reference here.  What these tests establish is that the
CUDA path computes that architecture: bf16 operands / fp32 accumulation and residual stream against fp32 on the CPU.
Tolerance (stated per test): ||got - ref|| / ||ref|| of the un-normalised features <= 2 % after two blocks, <= 3 % after all 32 (every
block rounds its GEMM and attention operands to bf16, ~0.4 % each), no element off by more than six times that, cosine >= 0.999; and
<= 0.8 % against the same restatement with its matrix-product operands rounded to bf16 (what separates a wrong kernel from rounding).
"""
import dataclasses

import pytest
import torch

from novic_b200 import default_decoder, synth
from novic_b200.encoder import EncoderDecoder, ImageEncoder, VitDims, synth_images, synth_vit_state_dict
from oracle import vit_oracle as vo

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _cfg(d: VitDims) -> vo.VitCfg:
    return vo.VitCfg(**dataclasses.asdict(d))


def _compare(d: VitDims, batch: int, chunk: int, rel_tol: float, seed: int = 7, tight_tol: float = 0.0):
    sd = synth_vit_state_dict(d, seed=seed)
    enc = ImageEncoder(d, images_per_chunk=chunk)
    enc.load_state_dict(sd, strict=True)
    enc = enc.to(DEV).eval()
    img = synth_images(batch, d, seed=seed + 4)
    with torch.inference_mode():
        ref = vo.encode_image(_cfg(d), sd, img)
        ref16 = vo.encode_image(_cfg(d), sd, img, bf16_operands=True) if tight_tol else None
        got = enc.encode_image(img.to(DEV)).cpu()
        got_n = enc.encode_image(img.to(DEV), normalize=True).cpu()
    assert got.shape == ref.shape == (batch, d.out_dim)
    rms = ref.pow(2).mean().sqrt().item()
    err = ((got - ref).norm() / ref.norm()).item()
    worst = (got - ref).abs().max().item() / rms
    cos = torch.nn.functional.cosine_similarity(got, ref, dim=1).min().item()
    print(f"{d.width}x{d.layers} layers, {d.tokens} tokens, batch {batch}: ||d|| / ||ref|| = {err:.4f}, max |d| / rms = {worst:.4f}, min cosine = {cos:.6f}")
    assert err <= rel_tol and worst <= 6 * rel_tol and cos >= 0.999
    if ref16 is not None:   # against the restatement that sees the same bf16-rounded operands: only accumulation order and the exp2 path differ
        err16 = ((got - ref16).norm() / ref16.norm()).item()
        print(f"    against the bf16-operand restatement: ||d|| / ||ref|| = {err16:.5f}")
        assert err16 <= tight_tol
    assert (got_n.norm(dim=1) - 1).abs().max().item() < 1e-5
    assert (got_n - torch.nn.functional.normalize(got, dim=1)).abs().max().item() < 1e-5
    return enc, img, got


def test_small_encoder_every_piece():
    """width 640 (8 heads of 80), 2 blocks, 26 tokens (one ragged query / key tile), batch 5 in chunks of 2 (ragged last chunk)."""
    _compare(VitDims(image_size=70, width=640, heads=8, layers=2, mlp_dim=2560), batch=5, chunk=2, rel_tol=0.02, tight_tol=0.008)


def test_full_width_two_blocks_730_tokens():
    """The real token count: 730 = 5 x 128 + 90 queries per (image, head), 11 x 64 + 26 keys; width 1280, 16 heads, mlp 5120."""
    _compare(VitDims(layers=2), batch=3, chunk=2, rel_tol=0.02, tight_tol=0.008)


def test_full_depth_encoder_feeds_the_decoder():
    """All 32 blocks on two images (the CPU restatement costs about 2 TFLOP), then images -> labels on the device: the embeddings the
    encoder hands over make the decoder generate exactly what it generates from the same embeddings passed in by the caller."""
    enc, img, feats = _compare(VitDims(), batch=2, chunk=2, rel_tol=0.03, seed=9)
    dims = synth.DecoderDims()
    dec = default_decoder(dims, synth.make_eos_friendly(synth.synth_state_dict(dims, seed=2, token_scale=0.25, jitter_norms=True), dims, beta=0.1)).to(DEV)
    both = EncoderDecoder(enc, dec)
    with torch.inference_mode():
        tok, pad, score = both.generate(img.to(DEV))
        e = torch.nn.functional.normalize(feats, dim=1).to(DEV)
        t2, p2, _, _, _, s2 = dec.generate(e, False, True, 1.0, 0.0, None, None, False)
    assert tok.shape[0] == 2 and tok.shape[1] == 1
    assert torch.equal(tok[:, 0], t2) and torch.equal(pad[:, 0], p2)
    assert (score[:, 0] - s2).abs().max().item() < 1e-3


def test_encoder_state_dict_keys_follow_open_clip():
    d = VitDims(image_size=70, width=640, heads=8, layers=2, mlp_dim=2560)
    keys = set(ImageEncoder(d).state_dict())
    assert {"visual.conv1.weight", "visual.class_embedding", "visual.positional_embedding", "visual.ln_pre.weight", "visual.ln_post.bias", "visual.proj",
            "visual.transformer.resblocks.1.attn.in_proj_weight", "visual.transformer.resblocks.1.attn.in_proj_bias",
            "visual.transformer.resblocks.0.attn.out_proj.weight", "visual.transformer.resblocks.0.mlp.c_fc.weight",
            "visual.transformer.resblocks.0.mlp.c_proj.bias", "visual.transformer.resblocks.0.ln_2.weight"} <= keys
    with pytest.raises(RuntimeError, match="no CPU path"):
        ImageEncoder(d).encode_image(synth_images(1, d))
