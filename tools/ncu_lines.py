#!/usr/bin/env python
"""Per-source-line summary of an ncu report: python tools/ncu_lines.py rep.ncu-rep [kernel-regex] [top N]
(reads `ncu -i rep --page source --csv --print-source cuda,sass`; needs -lineinfo at compile time)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr, lines, seen_kernel = None, None, [], 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        pass
    elif r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[2] == "-" and r[0].isdigit():
        d = dict(zip(hdr[4:], r[4:]))
        try:
            lines.append((cur_file, int(r[0]), r[1].strip(), int(d["Instructions Executed"] or 0), int(d["# Samples"] or 0),
                          float(d["Avg. Threads Executed"] or 0), {k: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v)}))
        except (KeyError, ValueError):
            continue
# the report may hold several launches of the same kernel: lines repeat; keep the first occurrence per (file, line)
uniq = {}
for l in lines:
    uniq.setdefault((l[0], l[1]), l)
lines = list(uniq.values())
tot_i = sum(l[3] for l in lines) or 1
tot_s = sum(l[4] for l in lines) or 1
print(f"total warp-instructions {tot_i}, samples {tot_s}")
print("== by instructions executed ==")
for l in sorted(lines, key=lambda l: -l[3])[:top]:
    print(f"{100*l[3]/tot_i:5.1f}% inst {100*l[4]/tot_s:5.1f}% smp thr {l[5]:4.1f} {l[0]}:{l[1]:<4} {l[2][:100]}")
print("== by stall samples ==")
for l in sorted(lines, key=lambda l: -l[4])[:top]:
    st = sorted(l[6].items(), key=lambda kv: -kv[1])[:3]
    print(f"{100*l[4]/tot_s:5.1f}% smp {100*l[3]/tot_i:5.1f}% inst {l[0]}:{l[1]:<4} {l[2][:80]}  {st}")
