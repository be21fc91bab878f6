import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, numpy as np
from novic_b200 import synth, default_decoder
from tests.golden_util import Golden, gold_embed, weight_case
gold = Golden()
dims = synth.DecoderDims()
tag, name = "lively", "gflat"
gt = synth.synth_guide_targets(300, dims, seed=21, first_pool=0)
model = default_decoder(dims, weight_case(tag)).to("cuda:0")
with torch.inference_mode():
    collect = os.environ.get("COLLECT", "0") == "1"
    tok, pad, lg, ls, lb, score = model.generate(gold_embed().cuda(), collect, True, 0.8, 0.3, None, gt.cuda(), False)
    _, _, lg, _, _, _ = model.generate(gold_embed().cuda(), True, True, 0.8, 0.3, None, gt.cuda(), False)
tok, pad, score, lg = tok.cpu(), pad.cpu(), score.cpu(), lg.cpu()
rt, rs = gold[f"{tag}/{name}/greedy_p/tok"], gold[f"{tag}/{name}/greedy_p/score"]
for b in range(tok.shape[0]):
    if not torch.equal(tok[b, :rt.shape[1]], rt[b, :tok.shape[1]]):
        print("row", b, "mine", tok[b].tolist(), "ref", rt[b].tolist(), "scores", score[b].item(), rs[b].item())
        # first-step logits of both first tokens
        first_allowed = sorted(set(gt[:, 0].tolist()))
        l0 = lg[b, 0]
        best = max(first_allowed, key=lambda t: l0[t].item())
        print("   step-1 logits: mine", l0[tok[b, 0]].item(), "ref", l0[rt[b, 0]].item(), "best allowed by my logits", best, l0[best].item())
np.save(os.path.join(ROOT, "gpurun_out", "guided_dbg_tok.npy"), tok.numpy())
