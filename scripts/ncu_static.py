#!/usr/bin/env python
"""Static SASS footprint by source line from an `ncu --page source --print-source cuda,sass --csv` export:
which source lines the kernel's code bytes come from, next to how often they run. usage: ncu_static.py export.csv [top_n]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, line, hdr = None, None, None
stat, dyn, src = defaultdict(set), defaultdict(int), {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if not hdr or len(r) < 8:
        continue
    if r[0] != "":
        line = (cur, int(r[0])); src[line] = r[1].strip()[:80]; continue
    if r[2].startswith("0x"):
        stat[line].add(r[2])
        try:
            dyn[line] += int(r[7] or 0)
        except ValueError:
            pass
tot = sum(len(v) for v in stat.values())
print(f"static SASS instructions {tot} = {tot * 16 / 1024:.1f} KB")
pf = defaultdict(int)
for (f, l), v in stat.items():
    pf[f] += len(v)
for f, c in sorted(pf.items(), key=lambda kv: -kv[1]):
    print(f"  {c:6d} {100 * c / tot:5.1f}%  {f}")
cold = sum(len(v) for k, v in stat.items() if dyn[k] == 0)
print(f"never executed in this capture: {cold} instr = {cold * 16 / 1024:.1f} KB")
print("top static lines:")
for k, v in sorted(stat.items(), key=lambda kv: -len(kv[1]))[:top]:
    print(f"  {len(v):5d} static {dyn[k]:12d} dyn  {k[0]}:{k[1]}  {src.get(k, '')}")
