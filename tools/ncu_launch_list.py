#!/usr/bin/env python
"""Launch-list summary: python tools/ncu_launch_list.py launches.csv "command line" > profiles/xxx_launch_list_summary.txt
Input = `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv <command>`.
Per kernel name: launches, total and average duration, share of all profiled time (cold-cache, serialised: compare SHARES)."""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if r]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
col = {h: i for i, h in enumerate(rows[hdr])}
agg = defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= col["Metric Value"] or r[col["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[col["Metric Value"]].replace(",", ""))
    unit = r[col["Metric Unit"]]
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    name = r[col["Kernel Name"]].split("(")[0]
    agg[name][0] += 1
    agg[name][1] += ms
total = sum(v[1] for v in agg.values())
print(f"# ncu --metrics gpu__time_duration.sum --clock-control none {sys.argv[2] if len(sys.argv) > 2 else ''} (cold-cache, serialised: compare SHARES)")
for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:60s} launches={n:4d} total_ms={ms:10.3f} avg_ms={ms / n:9.4f} share={100 * ms / total:5.1f}%")
