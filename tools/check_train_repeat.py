#!/usr/bin/env python
"""Repeatability of the training step's gradients at the bench's batch size (1024 samples): the same forward + backward three times
(dropout off and on with a fixed seed) - the gradients may only differ by fp32 atomic-order noise.  Catches races the 24-sample parity
tests are too small to provoke.      python tools/check_train_repeat.py [--batch 1024]
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from novic_b200 import synth, default_decoder
ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=1024); args = ap.parse_args()
dev = torch.device("cuda", 0)
dims = synth.DecoderDims()
out = {}
for p_drop in (0.0, 0.1):
    model = default_decoder(dims, synth.synth_state_dict(dims, seed=1), input_dropout=p_drop, layer_dropout=p_drop).to(dev).train()
    embed = synth.synth_embeddings(args.batch, seed=100).to(dev)
    tgt, pad = synth.synth_targets(args.batch, dims, seed=200); tgt, pad = tgt.to(dev), pad.to(dev)
    runs = []
    for rep in range(3):
        torch.manual_seed(7)
        model.zero_grad(set_to_none=True)
        _, _, loss_sum, loss_basis, _ = model(embed, tgt, pad, None, True, True, False, None)
        (loss_sum / loss_basis).backward()
        torch.cuda.synchronize()
        runs.append(((loss_sum / loss_basis).item(), {k: p.grad.detach().double().clone() for k, p in model.named_parameters()}))
    worst = 0.0; worst_key = None
    for rep in (1, 2):
        for k, g0 in runs[0][1].items():
            rel = (runs[rep][1][k] - g0).norm().item() / max(g0.norm().item(), 1e-30)
            if rel > worst: worst, worst_key = rel, k
    out[f"p={p_drop}"] = {"loss": [r[0] for r in runs], "worst_rel_diff": worst, "worst_key": worst_key,
                          "grad_norm": sum(g.norm().item() ** 2 for g in runs[0][1].values()) ** 0.5}
print(json.dumps(out))
