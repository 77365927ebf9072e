#!/usr/bin/env python
"""One line per profiled launch of an ncu report (python tools/ncu_summary.py rep.ncu-rep > profiles/xxx.txt):
duration, DRAM bytes, hit rates, occupancy, issue utilisation, divergence and the top stall reasons."""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
KEYS = [("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("l1tex__t_sector_hit_rate.pct", "l1hit%"), ("lts__t_sector_hit_rate.pct", "l2hit%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"), ("sm__inst_executed.sum", "warp_inst"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"), ("lts__t_sectors_op_atom.sum", "l2_atom_sectors"), ("lts__t_sectors_op_red.sum", "l2_red_sectors")]
print(f"# {rep}: ncu --set full --clock-control none (per-launch numbers are cold-cache, serialised)")
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0]
    parts = [name]
    for k, short in KEYS:
        if k in col:
            parts.append(f"{short}={r[col[k]]}{units[col[k]] if units[col[k]] not in ('', '%') else ''}")
    stalls = []
    for h, i in col.items():
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") or (h.startswith("smsp__average_warp_latency_issue_stalled_") and h.endswith(".ratio")):
            try:
                stalls.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", "").replace("_per_issue_active.ratio", "").replace(".ratio", "")))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    parts.append("stalls(cycles/issue): " + ", ".join(f"{n}={v:.2f}" for v, n in stalls[:5]))
    print("  ".join(parts))
