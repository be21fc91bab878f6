#!/usr/bin/env python
"""The command ncu captures (tools/ncu_capture_r02.sh): three greedy decodes of the bench batch with direct launches (no CUDA graph, so that
every kernel is its own ncu record).  Launches per decode: 1 embed-prep, 1 prefix GEMM, 6 prefill layers, then 14 decode steps x 6 layers x
(QKV GEMM, attention stream, fused block kernel) + 15 x (logits GEMM, greedy selection) + finalize."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
os.environ["NOVIC_NO_GRAPHS"] = "1"
import torch
from novic_b200 import default_decoder, synth
dims = synth.DecoderDims()
model = default_decoder(dims, synth.synth_state_dict(dims, seed=1)).to("cuda:0")
embed = synth.synth_embeddings(4096, seed=1234).to("cuda:0")
with torch.inference_mode():
    for _ in range(3):
        tok, pad, _, _, _, score = model.generate(embed, False, True, 1.0, 0.0, None, None, False)
torch.cuda.synchronize()
print("decoded", tuple(tok.shape), float(score.mean()))
