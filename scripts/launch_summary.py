#!/usr/bin/env python
"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> markdown share table.
    python scripts/launch_summary.py gpurun_out/launches_c5.csv "title" > profiles/rNN_launches_x.md"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    name = r[ki].split("(")[0]
    agg[name][0] += 1
    agg[name][1] += float(r[vi].replace(",", "")) * scale.get(r[ui], 1e-6)
tot = sum(v[1] for v in agg.values())
print(f"# {sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]}\n")
print("`ncu --metrics gpu__time_duration.sum --clock-control none`: per-launch times are cold-cache and serialised -- compare SHARES.")
print(f"{sum(v[0] for v in agg.values())} launches, {tot:.1f} ms in kernels.  Raw CSV beside this file.\n")
print("| kernel | launches | total ms | share | avg ms |\n|---|---|---|---|---|")
for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name[:110]}` | {n} | {ms:.3f} | {100 * ms / tot:.1f}% | {ms / n:.4f} |")
