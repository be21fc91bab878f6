"""Median in-graph time of one B=4096 greedy decode under the current environment (NOVIC_* toggles)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from novic_b200 import default_decoder, synth
dims = synth.DecoderDims()
m = default_decoder(dims, synth.synth_state_dict(dims, seed=1)).to('cuda:0')
e = synth.synth_embeddings(4096, seed=1234).to('cuda:0')
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda:0')
with torch.inference_mode():
    for _ in range(4):
        out = m.generate(e, False, True, 1.0, 0.0, None, None, False)
    ts = []
    for _ in range(20):
        flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); out = m.generate(e, False, True, 1.0, 0.0, None, None, False); b.record()
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
ts.sort()
print('%s median %.3f ms min %.3f max %.3f  checksum %d' % (' '.join(sys.argv[1:]), ts[len(ts) // 2], ts[0], ts[-1], int(out[0].sum().item())))
