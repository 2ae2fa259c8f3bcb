#!/usr/bin/env python
"""tools/ncu_regions.py -- fold the per-line shares printed by tools/ncu_lines.py into the code regions of dmv_gather.cu.

    python tools/ncu_lines.py REPORT --func Li64ELi1ELi1 --top 1000 | python tools/ncu_regions.py
"""
import collections
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(ROOT, "vlgae_b200", "csrc", "dmv_gather.cu")).read().split("\n")


def find(s, start=0):
    return next(i + 1 for i, l in enumerate(src) if s in l and i >= start)


lin = find("__device__ bool lin_pass")
marks = [(1, "helpers"), (find("template <int NT, bool WITH_CT>"), "stage"),
         (find("__device__ __forceinline__ void offsets"), "offsets"),
         (find("__device__ void log_pass"), "log_pass (fallback)"), (lin, "lin: setup/convert"),
         (find("// ---------------- inside", lin), "lin: inside"),
         (find("// ---------------- outside (linear", lin), "lin: outside setup"),
         (find("// phase A'(w)", lin), "lin: A'"), (find("// phase B'(w)", lin), "lin: B'"),
         (find("// ---------------- outputs: alpha", lin), "lin: outputs+check"),
         (find("__device__ __forceinline__ void amax"), "max helpers"), (find("__device__ void max_pass"), "max: setup"),
         (find("// Fused width step (as in lin_pass)"), "max: sweep"), (find("// back-trace: breadth-first"), "backtrace"),
         (find("__host__ __device__ inline size_t log_bytes"), "kernel main")]
reg, regs = collections.Counter(), collections.Counter()
for l in sys.stdin:
    m = re.match(r"\s*([\d.]+)% instr\s+([\d.]+)% samples\s+L\s*(\S+):\s*(.*)", l)
    if not m:
        if l.startswith("total"):
            print(l.strip())
        continue
    pi, ps, ln, text = float(m.group(1)), float(m.group(2)), m.group(3), m.group(4)
    if text.startswith("('"):
        name = "[" + text.split("'")[1] + "]"
    elif ln == "?":
        name = "?"
    else:
        name = [n for s_, n in marks if s_ <= int(ln)][-1]
    reg[name] += pi
    regs[name] += ps
for k, v in sorted(reg.items(), key=lambda x: -x[1]):
    print(f"{v:5.1f}% instr {regs[k]:5.1f}% samples  {k}")
