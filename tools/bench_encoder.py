#!/usr/bin/env python
"""Image encoder + decoder throughput (BASELINE config #5 shape: random-init CLIP ViT-H/14-378 + default decoder on synthetic 378 x 378
images).  python tools/bench_encoder.py [--images 256] [--chunk 64] [--steps 2]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
ap = argparse.ArgumentParser(); ap.add_argument("--images", type=int, default=256); ap.add_argument("--chunk", type=int, default=64)
ap.add_argument("--steps", type=int, default=2); ap.add_argument("--layers", type=int, default=32); args = ap.parse_args()
from novic_b200 import default_decoder, synth
from novic_b200.encoder import EncoderDecoder, ImageEncoder, VitDims, synth_images, synth_vit_state_dict
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))); torch.cuda.set_device(dev)
d = VitDims(layers=args.layers)
enc = ImageEncoder(d, images_per_chunk=args.chunk)
enc.load_state_dict(synth_vit_state_dict(d, seed=7)); enc = enc.to(dev).eval()
dims = synth.DecoderDims()
dec = default_decoder(dims, synth.synth_state_dict(dims, seed=1)).to(dev)
both = EncoderDecoder(enc, dec)
img = synth_images(min(args.images, 64), d).to(dev)
img = img.repeat((args.images + img.shape[0] - 1) // img.shape[0], 1, 1, 1)[:args.images].contiguous()
T, W, L, M = d.tokens, d.width, d.layers, d.mlp_dim
flops_img = 2 * T * L * (4 * W * W + 2 * W * M) + 4 * T * T * W * L + 2 * (T - 1) * 588 * W
def timed(fn, n):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): out = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n, out
with torch.inference_mode():
    enc_ms, e = timed(lambda: enc.encode_image(img, normalize=True), args.steps)
    all_ms, out = timed(lambda: both.generate(img), args.steps)
print(json.dumps({"images": args.images, "chunk": args.chunk, "encoder_ms": enc_ms, "encoder_images_per_s": args.images / enc_ms * 1e3,
                  "encoder_tflops": flops_img * args.images / enc_ms / 1e9, "gflop_per_image": flops_img / 1e9,
                  "end_to_end_ms": all_ms, "end_to_end_images_per_s": args.images / all_ms * 1e3, "labels_shape": list(out[0].shape)}))
