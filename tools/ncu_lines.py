#!/usr/bin/env python
"""tools/ncu_lines.py -- per-source-line instruction / stall-sample shares of one kernel in an .ncu-rep.

    python tools/ncu_lines.py REPORT.ncu-rep --func Li64ELi1ELi1 [--src vlgae_b200/csrc/dmv_gather.cu] [--top 40]

ncu's CSV export of the source page is SASS-only; this joins it with the line table of the in-tree library
(cuobjdump -xelf + nvdisasm -g) so that the counts can be read against the CUDA source.  Runs without a GPU.
"""
import argparse
import collections
import csv
import glob
import io
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_table(lib, func, cu_name):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, capture_output=True)
    stem = os.path.splitext(os.path.basename(cu_name))[0]
    cubin = [f for f in glob.glob(os.path.join(tmp, "*.cubin")) if os.path.basename(f).startswith(stem + ".")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
    start = next(i for i, l in enumerate(dis) if ".text." in l and func in l)
    table, cur = {}, None
    for l in dis[start + 1:]:
        if l.startswith(".section") and ".text" in l and func not in l:
            break
        m = re.search(r'//## File "(.*?)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            table[int(m.group(1), 16)] = (cur, m.group(2))
    return table


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--func", required=True, help="substring of the mangled kernel name (template arguments)")
    ap.add_argument("--kernel", default=None, help="ncu -k filter (regex:...)")
    ap.add_argument("--skip", type=int, default=0, help="index of the launch inside the report (ncu --launch-skip)")
    ap.add_argument("--src", default="vlgae_b200/csrc/dmv_gather.cu")
    ap.add_argument("--lib", default=os.path.join(ROOT, "vlgae_b200", "libvlgae_b200.so"))
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--sass", action="store_true", help="also list the hottest SASS instructions")
    args = ap.parse_args()
    cmd = ["ncu", "-i", args.report, "--page", "source", "--csv"]
    if args.kernel:
        cmd += ["-k", args.kernel]
    if args.skip:
        cmd += ["--launch-skip", str(args.skip), "--launch-count", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    h = rows[hi]
    ci, si = h.index("Instructions Executed"), h.index("# Samples")
    stalls = [i for i, k in enumerate(h) if k.startswith("stall_")]
    table = line_table(args.lib, args.func, args.src)
    src = open(os.path.join(ROOT, args.src)).read().split("\n")
    base = os.path.basename(args.src)
    by_i, by_s, st = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
    sass = []
    ti = ts = 0
    base_addr = None
    for r in rows[hi + 1:]:
        try:
            a, n, s = int(r[0], 16), int(r[ci]), int(r[si])
            if base_addr is None:
                base_addr = a  # the report lists absolute addresses: rebase on the kernel's first instruction
            a -= base_addr
        except (ValueError, IndexError):
            if len(r) > ci and r[0] == "Address":
                break  # next kernel
            continue
        key = table.get(a, (None, ""))[0]
        by_i[key] += n
        by_s[key] += s
        ti += n
        ts += s
        for k in stalls:
            try:
                st[key][h[k]] += int(r[k])
            except ValueError:
                pass
        sass.append((s, n, a, table.get(a, (None, r[1]))[1], key))
    print(f"total warp-instructions {ti}, stall samples {ts}")
    for key, n in sorted(by_i.items(), key=lambda x: -x[1])[:args.top]:
        text = src[key[1] - 1].strip()[:90] if key and key[0] == base else str(key)
        top = ", ".join(f"{k[6:]} {v / max(by_s[key], 1) * 100:.0f}%" for k, v in st[key].most_common(3))
        print(f"{n / ti * 100:5.1f}% instr {by_s[key] / max(ts, 1) * 100:5.1f}% samples  L{key[1] if key else '?':>4}: {text}   [{top}]")
    if args.sass:
        print("---- hottest SASS by samples")
        for s, n, a, t, key in sorted(sass, reverse=True)[:args.top]:
            print(f"{s / max(ts, 1) * 100:5.1f}% {n:10d}  {a:05x} L{key[1] if key else '?':>4} {t[:90]}")


if __name__ == "__main__":
    main()
