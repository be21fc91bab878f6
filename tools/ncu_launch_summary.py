#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (tools/ncu_capture_r02.sh writes gpurun_out/launches_<tag>.csv):
    python tools/ncu_launch_summary.py gpurun_out/launches_r02c.csv > profiles/r02c_launches_summary.txt
Times under ncu are cold-cache and serialised: compare the SHARES with bench.py's in-graph class times, not the absolutes."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
cols = rows[h]
ki, vi, ui = cols.index("Kernel Name"), cols.index("Metric Value"), cols.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r[ui], 1.0)
    k = re.sub(r"\(.*", "", r[ki])
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v for _, v in agg.values())
print(f"ncu --metrics gpu__time_duration.sum --clock-control none: {sys.argv[1]} (tools/ncu_target.py, direct launches, cold-cache, serialised); unit ns; total {tot}")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v:12.1f} {100 * v / tot:5.1f}%  x{n:4d}  {k[:120]}")
