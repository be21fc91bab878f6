"""Print the headline and the three largest kernel classes of every gpurun_out/b_wsd_*.json (tuning aid for NOVIC_QKV_WS_DIV sweeps)."""
import glob, json
for f in sorted(glob.glob("gpurun_out/b_wsd_*.json")):
    try:
        d = [json.loads(l) for l in open(f) if l.startswith("{")][-1]
        print(f, round(d["ms_per_step"], 4), round(d["value"]), {k: round(v["ms_per_step"], 4) for k, v in d["roofline"]["kernels"].items() if k in ("qkv_gemm", "attention", "block_outproj_ffn")})
    except Exception as e:
        print(f, "ERR", e)
