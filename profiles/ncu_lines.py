#!/usr/bin/env python
"""Summarise an ncu report per CUDA source line: python profiles/ncu_lines.py report.ncu-rep [top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
data = []
for r in rows:
    if len(r) > 6 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) > 7 and r[0].isdigit():
        d = dict(zip(hdr[4:], r[4:]))
        try:
            data.append((int(d["# Samples"]), int(d["Instructions Executed"]), int(r[0]), r[1].strip()[:100], d))
        except (KeyError, ValueError):
            pass
tot = sum(d[0] for d in data)
toti = sum(d[1] for d in data)
print(f"total samples {tot}  total warp-instructions {toti}")
stall_keys = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
for s, ins, line, src, d in sorted(data, key=lambda r: r[:3], reverse=True)[:top]:
    stalls = sorted(((int(d.get(k, 0) or 0), k[6:]) for k in stall_keys), reverse=True)[:3]
    st = " ".join(f"{k}:{v}" for v, k in stalls if v)
    print(f"{100.0 * s / max(tot, 1):5.1f}%  inst {ins:8d}  L{line:<4d} {src}\n        {st}")
