import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from novic_b200 import synth, default_decoder
from oracle import novic_oracle as orc
sys.path.insert(0, os.path.join(ROOT, "tests"))
from tests.golden_util import weight_case, gold_embed
tag = sys.argv[1] if len(sys.argv) > 1 else "eos"
H = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sd = weight_case(tag); cfg = orc.cfg_from_state_dict(sd)
model = default_decoder(synth.DecoderDims(), sd).to("cuda:0")
e = gold_embed()
with torch.inference_mode():
    o = orc.generate_beam(cfg, sd, e, H, 1.0, 0.0)
    t, p, s = [x.cpu() for x in model.generate_beam(e.cuda(), H, 1.0, 0.0, None, False, 0.0, None, False)]
bad = ((s - o["score"]).abs() > 0.5).any(dim=1).nonzero().flatten().tolist()
print("ATTN_V1", os.environ.get("NOVIC_ATTN_V1"), "bad rows", bad)
for b in bad[:3]:
    print("row", b)
    print(" ours  tok", t[b].tolist(), "score", [round(v, 2) for v in s[b].tolist()])
    print(" ours  pad", p[b].int().tolist())
    print(" oracle tok", o["target"][b].tolist(), "score", [round(v, 2) for v in o["score"][b].tolist()])
    print(" oracle pad", o["padding"][b].int().tolist())
