"""Top stall-sample SASS lines of a .ncu-rep (source page)."""
import csv, subprocess, sys
path = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out[1:]))
hdr = rows[0]; si = hdr.index("Warp Stall Sampling (All Samples)"); src = hdr.index("Source")
data = [(int(r[si] or 0), i, r[src]) for i, r in enumerate(rows[1:])]
tot = sum(d[0] for d in data) or 1
print("total samples", tot)
for s, i, t in sorted(data, reverse=True)[:n]:
    print(f"{s:6d} {100*s/tot:5.1f}%  #{i:4d}  {t.strip()[:110]}")
