#!/usr/bin/env python
"""generate_all throughput (teacher-forced scoring of every guide noun for every embedding, embedding_decoder.py:986-1079).
    python tools/bench_all.py [--embeds 64] [--nouns 3000] [--renorm]
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from novic_b200 import synth, default_decoder
ap = argparse.ArgumentParser(); ap.add_argument("--embeds", type=int, default=64); ap.add_argument("--nouns", type=int, default=3000)
ap.add_argument("--renorm", action="store_true"); args = ap.parse_args()
dims = synth.DecoderDims()
model = default_decoder(dims, synth.synth_state_dict(dims, seed=1)).to("cuda:0")
embed = synth.synth_embeddings(args.embeds, seed=1234).to("cuda:0")
gt = synth.synth_guide_targets(args.nouns, dims, seed=33, first_pool=200).to("cuda:0")
with torch.inference_mode():
    pre = model.precompute_generate_all(0.0, None, False, 0.0, gt, args.renorm)
    model.generate_all(embed[:4], 10, 1.0, 0.0, None, False, 0.0, gt, args.renorm, precompute=pre)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    tok, pad, score = model.generate_all(embed, 10, 1.0, 0.0, None, False, 0.0, gt, args.renorm, precompute=pre)
    b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
C = pre[0].shape[1]
print(json.dumps({"metric": "generate_all", "embeddings": args.embeds, "nouns": args.nouns, "positions_per_noun": C, "guide_renorm": args.renorm, "ms": ms,
                  "labels_per_s": args.embeds / (ms * 1e-3), "sequences_per_s": args.embeds * args.nouns / (ms * 1e-3)}))
