#!/usr/bin/env python
"""Section metrics ("details" page) of every kernel in an .ncu-rep as plain text:  python tools/ncu_details.py REPORT.ncu-rep"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "details", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
ki, bi, gi, si, mi, ui, vi = (hdr.index(k) for k in ("Kernel Name", "Block Size", "Grid Size", "Section Name", "Metric Name", "Metric Unit", "Metric Value"))
last = None
for r in rows[1:]:
    if len(r) <= vi or not r[mi]:
        continue
    key = (r[0], r[ki])
    if key != last:
        print(f"\n== {r[ki]}   block {r[bi]} grid {r[gi]}")
        last, sec = key, None
    if r[si] != sec:
        sec = r[si]
        print(f"  -- {sec}")
    print(f"     {r[mi]:58s} {r[vi]} {r[ui]}")
