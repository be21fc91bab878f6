#!/usr/bin/env python
"""profiles/traffic.json from the `ncu --set full` captures of tools/ncu_capture_r02.sh (run here, where ncu can read reports):
    python tools/ncu_traffic.py <tag> [dir]
Per kernel class: DRAM bytes of the captured launch (dram__bytes_read.sum + dram__bytes_write.sum), that launch's algorithmic bytes
(DESIGN.md section 5), their ratio, DRAM throughput %, tensor-pipe % of elapsed cycles, duration under ncu.  bench.py attaches these to its
roofline entries (`traffic`) and labels a class latency-bound when neither is near its roof."""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
src = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out")
B, E, V = 4096, 512, 6912
ALGO = {   # algorithmic HBM bytes of the captured launch (B = 4096, the decode step whose query sees 16 keys)
    "attention": B * (16 * 2 * E * 2 + 2 * E * 2),                 # 16 K rows + 16 V rows of 1 KB, q in, out
    "qkv_gemm": B * (2 * E + 3 * 2 * E),                           # LayerNorm rows in; q, k, v rows out (weights L2-resident)
    "block_outproj_ffn": B * (2 * E + 4 * E + 4 * E + 2 * E),      # attention rows + fp32 residual in; residual + next LayerNorm rows out
    "logits_gemm": B * (2 * E + -(-V // 64) * 32),                 # final rows in, one 32-byte partial per 64 vocabulary columns out
    "select": B * (-(-V // 64) * 32 + E * (4 + 4 + 2)),            # partials in; gathered embedding row in, residual + LayerNorm rows out
}


def metrics(path):
    out = open(path).read() if path.endswith(".csv") else subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(out.splitlines()) if len(r) > 10]
    hdr, vals = rows[0], rows[2]
    def get(name):
        for h, v in zip(hdr, vals):
            if h == name:
                return float(v.replace(",", ""))
        return None
    return {"kernel": vals[hdr.index("Kernel Name")][:100], "dram_read": get("dram__bytes_read.sum"), "dram_write": get("dram__bytes_write.sum"),
            "units": {h: rows[1][i] for i, h in enumerate(hdr) if h.startswith("dram__bytes")},
            "dram_pct": get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "tensor_pct": get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed") or get("sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed"),
            "duration": get("gpu__time_duration.sum"), "duration_unit": rows[1][hdr.index("gpu__time_duration.sum")],
            "warps_pct": get("sm__warps_active.avg.pct_of_peak_sustained_active")}


def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


out_path = os.path.join(ROOT, "profiles", "traffic.json")
res = json.load(open(out_path)) if os.path.isfile(out_path) else {}      # classes without a capture under this tag keep their entry
res["_comment"] =  "per kernel class: DRAM traffic of ONE launch from an ncu --set full capture (dram__bytes_read.sum + dram__bytes_write.sum; decode step with 16 visible keys, B = 4096, cold L2, serialised) beside that launch's algorithmic bytes; written by tools/ncu_traffic.py"
for name, algo in ALGO.items():
    path = os.path.join(src, f"{tag}_ncu_{name}_raw.csv")
    if not os.path.isfile(path):
        path = os.path.join(src, f"prof_{tag}_{name}.ncu-rep")
    if not os.path.isfile(path):
        continue
    m = metrics(path)
    rd = to_bytes(m["dram_read"], m["units"].get("dram__bytes_read.sum", "byte"))
    wr = to_bytes(m["dram_write"], m["units"].get("dram__bytes_write.sum", "byte"))
    res[name] = {"dram_bytes": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr, "algorithmic_bytes": algo, "ratio_to_algorithmic": (rd + wr) / algo,
                 "dram_pct": m["dram_pct"], "tensor_pipe_pct": m["tensor_pct"], "warps_active_pct": m["warps_pct"],
                 "duration": m["duration"], "duration_unit": m["duration_unit"], "kernel": m["kernel"], "capture": f"profiles/{tag}_ncu_{name}_raw.csv"}
    print(name, json.dumps(res[name]))
json.dump(res, open(out_path, "w"), indent=1)
