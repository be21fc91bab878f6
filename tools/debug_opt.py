import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from novic_b200 import synth, default_decoder
from tests.golden_util import weight_case
dims = synth.DecoderDims()
model = default_decoder(dims, weight_case("lively"), input_dropout=0.0, layer_dropout=0.0).to("cuda:0").train()
decay = [p for p in model.parameters() if p.dim() >= 2]; no_decay = [p for p in model.parameters() if p.dim() < 2]
opt = torch.optim.AdamW([{'params': no_decay, 'weight_decay': 0.0}, {'params': decay, 'weight_decay': 0.1}], lr=1.5e-3, betas=(0.9, 0.95), fused=True)
embed = synth.synth_embeddings(64, seed=3).cuda(); tgt, pad = synth.synth_targets(64, dims, seed=4); tgt, pad = tgt.cuda(), pad.cuda()
for i in range(8):
    opt.zero_grad(set_to_none=True)
    v0 = model.logits_linear.weight._version
    _, _, ls, lb, cor = model(embed, tgt, pad, None, True, True, False, None)
    (ls / lb).backward()
    norm = torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0, error_if_nonfinite=True)
    opt.step()
    print(i, (ls / lb).item(), "gradnorm", norm.item(), "correct", int(cor.sum()), "version", v0, model.logits_linear.weight._version)
