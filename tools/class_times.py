"""In-graph device time per kernel class (class-only replay, novic_debug_keep_classes) for any generation mode.
    python tools/class_times.py greedy|beam [topk] [guided_nouns] [batch]
"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from novic_b200 import _abi, default_decoder, synth
mode = sys.argv[1] if len(sys.argv) > 1 else "greedy"
topk = int(sys.argv[2]) if len(sys.argv) > 2 else 3
nguide = int(sys.argv[3]) if len(sys.argv) > 3 else 0
B = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
dims = synth.DecoderDims()
m = default_decoder(dims, synth.synth_state_dict(dims, seed=1)).to('cuda:0')
e = synth.synth_embeddings(B, seed=1234).to('cuda:0')
gt = synth.synth_guide_targets(nguide, dims, seed=33, first_pool=200).to('cuda:0') if nguide else None
lib = _abi.lib()
def run():
    if mode == "greedy":
        m.generate(e, False, True, 1.0, 0.0, None, gt, False)
    else:
        m.generate_beam(e, topk, 1.0, 0.0, None, False, 0.0, gt, False)
def timed(n=5):
    ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]
with torch.inference_mode():
    run(); run()
    h = m._state(e.device)["handle"]
    full = timed()
    _abi.check(lib.novic_debug_keep_classes(h, 1 << 31)); run(); base = timed()
    print(f"{mode} topk={topk} guided={nguide} B={B}: full {full:.3f} ms, empty graph {base:.3f} ms")
    tot = 0.0
    for i, name in enumerate(_abi.KERNEL_CLASSES):
        _abi.check(lib.novic_debug_keep_classes(h, 1 << i)); run()
        t = timed() - base
        if t > 0.003:
            print(f"   {name:14s} {t:8.3f} ms  {100 * t / full:5.1f} %")
            tot += t
    print(f"   sum of classes {tot:.3f} ms")
    _abi.check(lib.novic_debug_keep_classes(h, 0))
