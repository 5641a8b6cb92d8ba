#!/usr/bin/env python3
"""Aggregate an ncu source page (cuda,sass view) per source file and per source line.
usage: profiles/ncu_lines.py report.ncu-rep [kernel-id filter e.g. :::2] [top N]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
kid = sys.argv[2] if len(sys.argv) > 2 else ":::1"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id", kid],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, data = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur, hdr = r[1], None
    elif len(r) > 8 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0] != "":
        d = {"file": cur.split("/")[-1], "line": r[0], "src": r[1]}
        for k in ("# Samples", "Instructions Executed", "Thread Instructions Executed"):
            i = hdr.index(k)
            d[k] = float(r[i] or 0)
        data.append(d)
tot = sum(d["Instructions Executed"] for d in data) or 1
tots = sum(d["# Samples"] for d in data) or 1
byfile = collections.defaultdict(lambda: [0, 0, 0])
for d in data:
    b = byfile[d["file"]]
    b[0] += d["Instructions Executed"]; b[1] += d["Thread Instructions Executed"]; b[2] += d["# Samples"]
print(f"total warp instructions {tot/1e6:.1f} M, samples {tots:.0f}")
for k, b in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:22s} inst {100*b[0]/tot:5.1f}%  samples {100*b[2]/tots:5.1f}%  threads/inst {b[1]/max(b[0],1):5.1f}")
print("top lines by stall samples:")
for d in sorted(data, key=lambda d: -d["# Samples"])[:top]:
    print(f"  {d['file']:18s}:{d['line']:>4s} smp {100*d['# Samples']/tots:4.1f}% inst {100*d['Instructions Executed']/tot:4.1f}% "
          f"thr {d['Thread Instructions Executed']/max(d['Instructions Executed'],1):4.1f}  {d['src'].strip()[:80]}")
