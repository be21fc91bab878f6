import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from novic_b200 import synth, default_decoder
from oracle import novic_oracle as orc
from tests.golden_util import weight_case
L = int(sys.argv[1]) if len(sys.argv) > 1 else 6
B = int(sys.argv[2]) if len(sys.argv) > 2 else 24
use_pad = (sys.argv[3] == "1") if len(sys.argv) > 3 else True
dims = synth.DecoderDims(num_layers=L)
sd = synth.make_eos_friendly(synth.synth_state_dict(dims, seed=2, token_scale=0.25, jitter_norms=True), dims, beta=0.1)
cfg = orc.cfg_from_state_dict(sd)
embed = synth.synth_embeddings(B, seed=21)
tgt, pad = synth.synth_targets(B, dims, seed=5)
if not use_pad: pad = None
leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k != "causality_mask"}
_, ls, lb, cor = orc.forward_loss(cfg, leaf, embed, tgt, pad, None)
ls.backward()
model = default_decoder(dims, sd, num_layers=L, input_dropout=0.0, layer_dropout=0.0).to("cuda:0").train()
out = model(embed.cuda(), tgt.cuda(), None if pad is None else pad.cuda(), None, True, True, False, None)
print(f"L={L} B={B} pad={use_pad}: loss {out[2].item():.4f} vs {ls.item():.4f}  basis {float(out[3])} vs {float(lb)}  correct agree {(out[4].cpu()==cor).float().mean().item():.4f}")
out[2].backward()
got = dict(model.named_parameters())
for k, v in leaf.items():
    g = got[k].grad.detach().cpu().double(); r = v.grad.double()
    rel = (g - r).norm().item() / max(r.norm().item(), 1e-12)
    cos = torch.nn.functional.cosine_similarity(g.flatten(), r.flatten(), dim=0).item()
    flag = "" if (rel < 0.05 and cos > 0.998) else "   <-- BAD"
    print(f"  {k:50s} |ref| {r.norm().item():10.4f} |got| {g.norm().item():10.4f} rel {rel:8.4f} cos {cos:8.5f}{flag}")
