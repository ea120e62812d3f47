"""Per-kernel totals of an ncu launch list (`--metrics gpu__time_duration.sum --csv`).
Usage: python tools/ncu_launch_summary.py launches.csv "<command that was profiled>" > summary.json"""
import csv, json, re, sys
from collections import defaultdict
rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = defaultdict(lambda: [0, 0.0])
for r in rd:
    if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    us = v / 1e3 if r[ui] in ("ns", "nsecond") else v * (1e3 if r[ui] in ("ms", "msecond") else 1.0)
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("ptts::", "").replace("(anonymous namespace)::", "")
    name = re.sub(r"^<unnamed>::", "", name)
    t = tot[name]
    t[0] += 1
    t[1] += us
total = sum(t[1] for t in tot.values())
out = {"command": sys.argv[2] if len(sys.argv) > 2 else "", "total_us": round(total, 1),
       "launches": sum(t[0] for t in tot.values()),
       "kernels": [{"kernel": k, "launches": t[0], "total_us": round(t[1], 1), "share": round(t[1] / total, 4)}
                   for k, t in sorted(tot.items(), key=lambda kv: -kv[1][1])]}
print(json.dumps(out, indent=1))
