"""Per-phase clock64 trace of the fused layer-stack kernel (CTA (0,1), first epilogue thread), decode step, M = 4096."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from novic_b200 import synth, default_decoder, _abi
lib = _abi.lib()
dims = synth.DecoderDims(num_layers=2)
model = default_decoder(dims, synth.synth_state_dict(dims, seed=1), num_layers=2).to("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
e = synth.synth_embeddings(B, seed=1234).cuda()
names = {0: "entry", 1: "setup done (barriers, TMEM)", 2: "griddepcontrol.wait passed", 3: "OUT: acc + stats", 4: "OUT: after CB1", 5: "OUT: LN rows stored",
         6: "OUT: after CB2", 7: "FFN: linear1 acc ready", 8: "FFN: hidden tile written", 9: "FFN: linear2 acc + stats", 10: "FFN: after CB3",
         11: "FFN: after CB4 (LN rows stored)", 12: "QKV: tile 0 ready", 13: "QKV: tile 1 ready", 14: "QKV: tile 2 ready", 15: "exit", 16: "QKV: epilogue done", 17: "ATTN: after CB (q/k/v visible)", 18: "ATTN: this warp done", 19: "ATTN: after CB (ao visible)"}
with torch.inference_mode():
    st = model._state(torch.device("cuda:0"))
    _abi.check(lib.novic_set_use_graphs(st["handle"], 0))
    model.generate(e, False, True, 1.0, 0.0, None, None, False)
    # instrumented launches with a CTA (0,1), L = 2: prefix(0) QKV0(1) [OUT0 FFN0 QKV1](2) [OUT1 FFN1](3) | step 1: QKV0(4) [OUT0 FFN0 QKV1](5) [OUT1 FFN1](6)
    fused = os.environ.get("NOVIC_STACK_ATTN", "1") != "0"
    plan = {4: "whole stack, L = 2 (decode step 1, 5 keys)", 11: "whole stack, L = 2 (decode step 8, 12 keys)", 17: "whole stack, L = 2 (decode step 14, 18 keys)"} if fused else {4: "QKV_0 only (decode)", 5: "OUT_0 FFN_0 QKV_1 (decode)", 6: "OUT_1 FFN_1 (decode)"}
    for target, label in plan.items():
        _abi.check(lib.novic_debug_trace(None, 1 + target))
        model.generate(e, False, True, 1.0, 0.0, None, None, False)
        buf = (C.c_int64 * 32)()
        _abi.check(lib.novic_debug_trace(buf, 0))
        t0 = buf[0]
        print(label)
        for i in sorted(names, key=lambda k: buf[k]):
            if buf[i]:
                print(f"   {names[i]:36s} +{buf[i] - t0:7d} cycles")
