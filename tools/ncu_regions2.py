#!/usr/bin/env python
"""Like ncu_regions.py but regions may name their file: python tools/ncu_regions2.py rep.ncu-rep file:name:lo-hi ...
Prints warp instructions, thread instructions, average active threads and stall samples per region; lines outside every
region are listed per file."""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
regions = []
for a in sys.argv[2:]:
    f, n, r = a.split(":")
    lo, hi = r.split("-")
    regions.append((f, n, int(lo), int(hi)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr = None, None
seen = set()
agg = defaultdict(lambda: [0, 0, 0])
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[2] == "-" and r[0].isdigit():
        key = (cur_file, int(r[0]))
        if key in seen:
            continue
        seen.add(key)
        d = dict(zip(hdr[4:], r[4:]))
        try:
            wi, ti, smp = int(d["Instructions Executed"] or 0), int(d["Thread Instructions Executed"] or 0), int(d["# Samples"] or 0)
        except ValueError:
            continue
        name = "(other) " + cur_file
        for f, n, lo, hi in regions:
            if f == cur_file and lo <= key[1] <= hi:
                name = n
                break
        a = agg[name]
        a[0] += wi; a[1] += ti; a[2] += smp
tw, tt, ts = (sum(a[i] for a in agg.values()) or 1 for i in range(3))
print(f"total: warp-inst {tw}  thread-inst {tt}  avg active {tt / tw:.1f}  samples {ts}")
for n, (wi, ti, smp) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{n:40s} warp-inst {100 * wi / tw:5.1f}%  thread-inst {100 * ti / tt:5.1f}%  active {ti / max(wi, 1):5.1f}  samples {100 * smp / ts:5.1f}%")
