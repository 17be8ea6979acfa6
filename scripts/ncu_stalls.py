#!/usr/bin/env python
"""Stall-reason totals and the top source lines per stall reason from an `ncu --page source --print-source cuda,sass --csv`
export. usage: ncu_stalls.py export.csv [reason=stall_no_inst] [top_n]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
reason = sys.argv[2] if len(sys.argv) > 2 else "stall_no_inst"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
cur, hdr = None, None
tot = defaultdict(int)
lines = []
sass_rows = defaultdict(int)
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1]; continue
    if r[0] == "Line No":
        hdr = r
        idx = {n: i for i, n in enumerate(hdr)}
        stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
        continue
    if not hdr or len(r) < len(hdr):
        continue
    if r[2] == "-":   # per-source-line summary row
        try:
            vals = {n: int(r[idx[n]] or 0) for n in stall_cols}
        except ValueError:
            continue
        for n, v in vals.items():
            tot[n] += v
        lines.append((vals.get(reason, 0), int(r[6] or 0), (cur or "").split("/")[-1], r[0], r[1].strip()[:90]))
    else:
        sass_rows[((cur or "").split("/")[-1], r[0])] += 1
allsamp = sum(tot.values()) or 1
print("stall totals (all samples):")
for n, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v:
        print(f"  {100 * v / allsamp:5.1f}%  {n}")
print(f"top lines by {reason}:")
rt = tot[reason] or 1
for v, s, f, ln, src in sorted(lines, reverse=True)[:top]:
    print(f"  {100 * v / rt:5.1f}% of {reason} ({v:6d} / {s:6d} samples) sass={sass_rows[(f, ln)]:5d} {f}:{ln}  {src}")
