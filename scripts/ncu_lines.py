#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export by CUDA source line.
usage: ncu_lines.py export.csv [top_n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) >= 9 and r[2] == "-":      # per-source-line summary row
        try:
            samp, ins, thr = int(r[6]), int(r[7]), int(r[8])
        except ValueError:
            continue
        out.append((ins, samp, thr, (cur or "").split("/")[-1], r[0], r[1].strip()[:100]))
tot = sum(o[0] for o in out) or 1
tots = sum(o[1] for o in out) or 1
print(f"total warp-inst {tot}  samples {tots}")
for o in sorted(out, reverse=True)[:top]:
    print(f"{100*o[0]/tot:5.1f}% inst {100*o[1]/tots:5.1f}% samp lanes={o[2]/max(o[0],1):4.1f} {o[3]}:{o[4]}  {o[5]}")
