"""In-graph cost of each kernel class by ablation: time the captured greedy decode (B = 4096) with the launches of one class
dropped (NOVIC_SKIP_CLASSES bit mask; results are garbage, timing is what matters).  full - ablated = what the class costs
inside the graph, with PDL overlap and warm L2 - unlike ncu's serialised cold-cache per-launch times.

    python tools/ablate.py            # prints one line per configuration
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLASSES = ["prep", "prefix", "qkv", "attn", "outproj", "ffn1", "ffn2", "logits", "select", "misc"]
CHILD = r"""
import sys, torch
sys.path.insert(0, %r)
from novic_b200 import default_decoder, synth
dims = synth.DecoderDims()
m = default_decoder(dims, synth.synth_state_dict(dims, seed=1)).to('cuda:0')
e = synth.synth_embeddings(4096, seed=1234).to('cuda:0')
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda:0')
with torch.inference_mode():
    for _ in range(4):
        m.generate(e, False, True, 1.0, 0.0, None, None, False)
    ts = []
    for _ in range(20):
        flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); m.generate(e, False, True, 1.0, 0.0, None, None, False); b.record()
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
ts.sort()
print('%%.3f %%.3f %%.3f' %% (ts[len(ts) // 2], ts[0], ts[-1]))
""" % ROOT


def run(mask, extra_env=None):
    env = dict(os.environ, NOVIC_SKIP_CLASSES=str(mask))
    env.update(extra_env or {})
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    return out.stdout.strip() or out.stderr.strip()[-300:]


if __name__ == "__main__":
    print("config                      median_ms min_ms max_ms")
    print(f"{'full':28s}{run(0)}")
    for name in ("qkv", "attn", "outproj", "ffn2", "logits", "select"):
        print(f"{'without ' + name:28s}{run(1 << CLASSES.index(name))}")
    allg = sum(1 << CLASSES.index(n) for n in ("qkv", "outproj", "ffn2", "logits"))
    print(f"{'without all GEMMs':28s}{run(allg)}")
    print(f"{'attention only':28s}{run(allg | 1 << CLASSES.index('select'))}")
    print(f"{'GEMMs only (no attn)':28s}{run(1 << CLASSES.index('attn'))}")
    for cfg in sys.argv[1:]:
        k, v = cfg.split("=")
        print(f"{'full ' + cfg:28s}{run(0, {k: v})}")
