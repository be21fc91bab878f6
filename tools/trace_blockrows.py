"""Per-phase clock64 trace of the row-owner block kernel (blockrows.cuh; CTA 0, decode step, 1-layer model)."""
import ctypes as C, os, sys
os.environ.setdefault("NOVIC_BLOCK_ROWS", "32")
os.environ.setdefault("NOVIC_BLOCK_ROWS64_MIN", "0")   # the prefix pass on the traced kernel too, so that the launch ordinals below hold
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from novic_b200 import synth, default_decoder, _abi
lib = _abi.lib()
dims = synth.DecoderDims(num_layers=1)
model = default_decoder(dims, synth.synth_state_dict(dims, seed=1), num_layers=1).to("cuda:0")
e = synth.synth_embeddings(int(os.environ.get("TRACE_ROWS", "4096")), seed=1234).cuda()
names = {0: "entry", 1: "epilogue: wait released", 2: "epilogue: acc0 added (4 tiles)", 3: "epilogue: LN2 rows written", 4: "epilogue: acc1 ready", 5: "epilogue: hidden written",
         6: "epilogue: acc2 added", 7: "epilogue: stores issued", 8: "exit", 22: "MMA: attention rows landed", 23: "MMA: LN2 rows ready", 24: "MMA: hidden rows ready",
         25: "epilogue: acc0 tile 0 ready", 26: "epilogue: acc2 tile 0 ready", 27: "epilogue: x stores issued", 28: "epilogue: final statistics", 29: "epilogue: xn tile staged"}
names.update({10 + i: f"MMA: weight request {i} landed" for i in range(12)})
with torch.inference_mode():
    st = model._state(torch.device("cuda:0"))
    _abi.check(lib.novic_set_use_graphs(st["handle"], 0))
    model.generate(e, False, True, 1.0, 0.0, None, None, False)
    # GEMM launches of a 1-layer decode: prefix, qkv, block, logits, then per step qkv, block, logits -> the second block launch is ordinal 5
    _abi.check(lib.novic_debug_trace(None, 1 + 5))
    model.generate(e, False, True, 1.0, 0.0, None, None, False)
    buf = (C.c_int64 * 32)()
    _abi.check(lib.novic_debug_trace(buf, 0))
    for i, n in sorted(names.items(), key=lambda t: buf[t[0]]):
        if buf[i]:
            print(f"   {n:36s} +{buf[i] - buf[0]:7d} cycles")
