#!/usr/bin/env python
"""The commands compute-sanitizer runs (tools/sanitize.sh; SURVEY.md section 5's sanitizer lane): small invocations that launch every
kernel class of the hot path once or twice - the mbarrier / TMA / DSMEM / tcgen05 protocols are the same at any batch size.
    python tools/sanitize_target.py decode    greedy + beam-3 + guided beam + teacher-forced forward on 48 embeddings (graph replay AND direct launches)
    python tools/sanitize_target.py train     two training steps (noise -> fwd + bwd -> fused clip + AdamW) on 48 samples, dropout on
    python tools/sanitize_target.py encoder   one ViT block stack pass on 2 images (NOVIC_SAN_VIT_LAYERS blocks, default 2)
Prints a checksum line; correctness against the oracle is the job of tests/ - here the sanitizer's report is the result."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
mode = sys.argv[1] if len(sys.argv) > 1 else "decode"
import torch
from novic_b200 import default_decoder, synth
dev = torch.device("cuda", 0)
dims = synth.DecoderDims()
B = int(os.environ.get("NOVIC_SAN_BATCH", "48"))

if mode == "decode":
    sd = synth.make_eos_friendly(synth.synth_state_dict(dims, seed=2, token_scale=0.25, jitter_norms=True), dims, beta=0.1)
    model = default_decoder(dims, sd).to(dev)
    embed = synth.synth_embeddings(B, seed=1234).to(dev)
    tgt, tpad = synth.synth_targets(B, dims, seed=5)
    guide, _ = synth.synth_targets(40, dims, seed=9)
    with torch.inference_mode():
        for rep in range(2):            # the second call replays the captured graph
            tok, pad, _, _, _, score = model.generate(embed, False, True, 1.0, 0.0, None, None, False)
        bt, bp, bs = model.generate_beam(embed, 3, 1.0, 0.0, None, False, 0.0, None, False)
        gt, gp, gs = model.generate_beam(embed, 3, 1.0, 0.0, None, False, 0.0, guide.to(dev), True)
        out = model(embed, tgt.to(dev), tpad.to(dev), None, True, True, False, None)
    torch.cuda.synchronize()
    print("decode ok", tuple(tok.shape), float(score.sum()), float(bs.sum()), float(gs.sum()), float(out[2]))
elif mode == "train":
    from novic_b200 import EmbeddingNoise
    from novic_b200.optim import FusedAdamW
    from novic_b200.dist import train_step
    torch.manual_seed(11)
    model = default_decoder(dims, synth.synth_state_dict(dims, seed=1)).to(dev).train()
    opt = FusedAdamW(model, lr=1.5e-3, betas=(0.9, 0.95), weight_decay=0.1)
    noise = EmbeddingNoise.create("GaussElemUniformAngle", 1024, 3.25, 45.0, 75.0, 0.0, 0.15)
    embed = synth.synth_embeddings(B, seed=100).to(dev)
    tgt, pad = synth.synth_targets(B, dims, seed=200)
    tgt, pad = tgt.to(dev), pad.to(dev)
    losses = []
    for _ in range(2):
        loss, ncorrect, ntok, norm = train_step(model, opt, embed.clone(), tgt, pad, None, noise=noise, gradient_clip=1.0)
        losses.append(float(loss))
    torch.cuda.synchronize()
    print("train ok", losses)
elif mode == "encoder":
    from novic_b200.encoder import ImageEncoder, VitDims
    vd = VitDims(layers=int(os.environ.get("NOVIC_SAN_VIT_LAYERS", "2")))
    enc = ImageEncoder(vd, images_per_chunk=2).to(dev)
    img = torch.randn(2, 3, vd.image_size, vd.image_size, device=dev)
    with torch.inference_mode():
        e = enc(img)
    torch.cuda.synchronize()
    print("encoder ok", tuple(e.shape), float(e.float().abs().sum()))
else:
    raise SystemExit(f"unknown mode {mode}")
