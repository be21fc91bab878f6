import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from novic_b200 import synth, default_decoder
from tests.golden_util import weight_case, gold_embed, Golden
g = Golden()
for tag in ("lively", "eos", "eosall"):
    model = default_decoder(synth.DecoderDims(), weight_case(tag)).to("cuda:0")
    e = gold_embed().cuda()
    for name, H, tau, alpha in (("b3", 3, 1.0, 0.0), ("b5", 5, 1.3, 0.6), ("b10", 10, 1.0, 0.0)):
        with torch.inference_mode():
            t, p, s = [x.cpu() for x in model.generate_beam(e, H, tau, alpha, None, False, 0.0, None, False)]
            gr = model.generate(e, False, True, tau, alpha, None, None, False)[5].cpu()
        gt, gs = g[f"{tag}/{name}/tok"], g[f"{tag}/{name}/score"]
        same = (t == gt).all(dim=2) if t.shape == gt.shape else torch.zeros(t.shape[:2], dtype=torch.bool)
        print(f"{tag} {name}: T {t.shape[2]} vs {gt.shape[2]} best-same {same[:,0].float().mean():.3f} all-same {same.float().mean():.3f} "
              f"mean best diff {(s[:,0].mean()-gs[:,0].mean()).item():+.4f} min(best-greedy) {(s[:,0]-gr).min().item():+.4f} max|d| same {(s-gs)[same].abs().max().item() if same.any() else 0:.4f}")
