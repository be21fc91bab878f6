"""Wall-clock time of greedy and beam-3 decodes (B from argv) with output checksums - to compare NOVIC_* switches for speed AND equal results."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from novic_b200 import synth, default_decoder
dims = synth.DecoderDims()
model = default_decoder(dims, synth.synth_state_dict(dims, seed=1)).to("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
embed = synth.synth_embeddings(B, seed=1234).cuda()
with torch.inference_mode():
    for _ in range(3): out = model.generate(embed, False, True, 1.0, 0.0, None, None, False)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): out = model.generate(embed, False, True, 1.0, 0.0, None, None, False)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print(f"chains={os.environ.get('NOVIC_CHAINS','default')} greedy B={B}: {dt*1e3:.2f} ms -> {B/dt:,.0f} labels/s  checksum tok={int(out[0].sum())} score={out[5].sum().item():.3f}", flush=True)
    for _ in range(2): ob = model.generate_beam(embed, 3, 1.0, 0.0, None, False, 0.0, None, False)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): ob = model.generate_beam(embed, 3, 1.0, 0.0, None, False, 0.0, None, False)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    print(f"   beam3: {dt*1e3:.2f} ms -> {B/dt:,.0f} labels/s checksum {int(ob[0].sum())} {ob[2].sum().item():.3f}", flush=True)
