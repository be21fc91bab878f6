timeout 300 python -m pytest tests/test_guided.py -m gpu -q 2>&1 | grep -E "^E  |passed|failed|FAILED" | head -20
timeout 300 python -m pytest tests -m gpu -x -q --deselect tests/test_guided.py 2>&1 | tail -2
python - <<'PY'
import sys, time, torch
sys.path.insert(0, '.')
from novic_b200 import synth, default_decoder
dims = synth.DecoderDims()
model = default_decoder(dims, synth.synth_state_dict(dims, seed=1)).to("cuda:0")
embed = synth.synth_embeddings(4096, seed=1234).cuda()
gt = synth.synth_guide_targets(43000, dims, seed=33, first_pool=3000).cuda()
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
with torch.inference_mode():
    print("beam10 unguided            %.2f ms" % t(lambda: model.generate_beam(embed, 10, 1.0, 0.0, None, False, 0.0, None, False)))
    print("beam10 guided (43k nouns)  %.2f ms" % t(lambda: model.generate_beam(embed, 10, 1.0, 0.0, None, False, 0.0, gt, False)))
    print("beam10 guided renorm       %.2f ms" % t(lambda: model.generate_beam(embed, 10, 1.0, 0.0, None, False, 0.0, gt, True)))
    print("greedy unguided            %.2f ms" % t(lambda: model.generate(embed, False, True, 1.0, 0.0, None, None, False)))
    print("greedy guided              %.2f ms" % t(lambda: model.generate(embed, False, True, 1.0, 0.0, None, gt, False)))
PY
