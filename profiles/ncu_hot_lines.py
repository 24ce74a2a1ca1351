#!/usr/bin/env python
"""Summarise an `ncu --page source --print-source cuda,sass --csv` dump: executed instructions and
stall samples per CUDA source line.  usage: ncu_hot_lines.py dump.csv [top_n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = []
for r in rows:
    if len(r) >= 8 and r[0].isdigit() and r[2] == "-":
        try:
            out.append((int(r[7]), int(r[6]), int(r[0]), r[1].strip()[:120]))
        except ValueError:
            pass
tot = sum(o[0] for o in out) or 1
smp = sum(o[1] for o in out) or 1
print(f"total warp-instructions {tot}, samples {smp}")
for n, s, ln, src in sorted(out, reverse=True)[:top]:
    print(f"{n / tot * 100:5.1f}% inst {s / smp * 100:5.1f}% smp  L{ln}: {src}")
