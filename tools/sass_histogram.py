#!/usr/bin/env python
"""Opcode histogram of the SASS in a library: python tools/sass_histogram.py tg_b200/libtgb200.so > profiles/rNN_sass_histogram.txt
(cuobjdump -sass; per kernel: instruction count and the opcodes that tell which hardware paths the code uses)."""
import collections
import re
import subprocess
import sys

out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
kern, hist = None, collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
tot = collections.Counter()
for c in hist.values():
    tot.update(c)
WATCH = ["UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "HMMA", "UTCHMMA", "LDTM", "STTM", "REDUX", "VOTE", "MATCH", "SHFL", "ATOMG", "ATOMS", "ATOM", "RED", "LDS", "STS",
         "LDG", "STG", "LDL", "STL", "BAR", "MEMBAR", "NANOSLEEP", "MUFU", "FCHK", "POPC", "FLO", "BSSY", "BSYNC", "WARPSYNC"]
print(f"# cuobjdump -sass {sys.argv[1]}: {len(hist)} kernels, {sum(tot.values())} SASS instructions")
print("# Blackwell / Hopper-only paths (TMA = UTMALDG/UTMASTG/UBLKCP, cp.async = LDGSTS, tensor cores = HMMA/UTC*MMA, TMEM = LDTM/STTM): "
      + ", ".join(f"{op}={tot.get(op, 0)}" for op in ("UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "HMMA", "UTCHMMA", "LDTM", "STTM")))
print("# whole library: " + ", ".join(f"{op}={n}" for op, n in tot.most_common(30)))
for k in sorted(hist, key=lambda k: -sum(hist[k].values())):
    c = hist[k]
    print(f"{k:60s} n={sum(c.values()):6d}  " + " ".join(f"{op}={c[op]}" for op in WATCH if c.get(op)))
