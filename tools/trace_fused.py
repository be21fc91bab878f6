"""Per-phase clock64 trace of the fused out-proj + feed-forward cluster kernel (CTA (0,1), decode step, 1-layer model)."""
import ctypes as C, os, sys
os.environ.setdefault("NOVIC_BLOCK_ROWS", "128")   # this tool traces the 128-row cluster kernel (the default block kernel has tools/trace_blockrows.py)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from novic_b200 import synth, default_decoder, _abi
lib = _abi.lib()
dims = synth.DecoderDims(num_layers=1)
model = default_decoder(dims, synth.synth_state_dict(dims, seed=1), num_layers=1).to("cuda:0")
e = synth.synth_embeddings(4096, seed=1234).cuda()
names = ["entry", "wait released", "acc0 ready", "stats A written", "after barrier 1", "LN2 rows broadcast", "after barrier 2", "acc1 ready", "hidden broadcast",
         "after barrier 3", "acc2 ready", "stats C written", "after barrier 4", "stores issued", "exit barrier passed", "peer LN2 slices landed (MMA thread)"]
with torch.inference_mode():
    st = model._state(torch.device("cuda:0"))
    _abi.check(lib.novic_set_use_graphs(st["handle"], 0))
    model.generate(e, False, True, 1.0, 0.0, None, None, False)
    # GEMM launches of a 1-layer decode: prefix, qkv, fused, logits, then per step qkv, fused, logits -> the second fused launch is ordinal 5
    _abi.check(lib.novic_debug_trace(None, 1 + 5))
    model.generate(e, False, True, 1.0, 0.0, None, None, False)
    buf = (C.c_int64 * 32)()
    _abi.check(lib.novic_debug_trace(buf, 0))
    for i, n in sorted(enumerate(names), key=lambda t: buf[t[0]]):
        if buf[i]:
            print(f"   {n:24s} +{buf[i] - buf[0]:7d} cycles")
