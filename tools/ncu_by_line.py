#!/usr/bin/env python3
"""Aggregate the stall samples of an ncu report (--import-source on, -lineinfo) by source line.
    python tools/ncu_by_line.py <report.ncu-rep> [top N]"""
import collections
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, cols, ix = "", None, None
lines = []
cur = None   # [samples, instructions, file, line, source, stalls] of the source line whose SASS rows follow
for r in rows:
    if not r:
        continue
    if r[0] in ("File Path", "File Name"):
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        cols = r
        ix = {}
        for k, c in enumerate(cols):
            ix.setdefault(c, k)
        continue
    if cols is None:
        continue
    if r[0]:          # a source row: only its line number and text are used -- quotes or commas in the source text shift
        cur = [0, 0, fname, r[0], r[1].strip()[:70], collections.Counter()]   # its other columns; the SASS rows below carry the numbers
        lines.append(cur)
        continue
    if cur is None or len(r) < len(cols):
        continue
    try:
        cur[0] += int(r[ix["# Samples"]] or 0)
        cur[1] += int(r[ix["Instructions Executed"]] or 0)
    except ValueError:
        continue
    for k, c in enumerate(cols):
        if c.startswith("stall_") and "Not Issued" not in c and r[k] not in ("", "0"):
            cur[5][c[6:]] += int(float(r[k]))
lines = [tuple(l) for l in lines if l[0] > 0]
total = sum(l[0] for l in lines)
byfile = collections.Counter()
for l in lines:
    byfile[l[2]] += l[0]
print(f"total samples {total};  by file: " + ", ".join(f"{f} {100.0 * s / total:.1f}%" for f, s in byfile.most_common()))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for s, ninst, f, ln, src, st in sorted(lines, reverse=True)[:top]:
    sts = ", ".join(f"{c} {v}" for c, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100.0 * s / total:5.2f}% {ninst:10d} {f}:{ln:4s} {src:70s} | {sts}")
