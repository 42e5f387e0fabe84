#!/usr/bin/env python
"""Top SASS instructions of an .ncu-rep by warp-stall samples (source page).  usage: ncu_hot.py rep [N]"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
isrc, ismp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = []
tot = 0
for n, r in enumerate(rows[2:]):
    try:
        s = float(r[ismp]); tot += s
        data.append((s, n, r[isrc].strip()[:100], r[iex]))
    except Exception:
        pass
data.sort(reverse=True)
print("total samples", tot)
for s, n, src, ex in data[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{100*s/tot:5.1f}%  line {n:4d}  exec {ex:>9s}  {src}")
