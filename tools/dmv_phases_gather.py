#!/usr/bin/env python
"""tools/dmv_phases_gather.py -- cycle counts of the sentence-0 CTAs under the gather schedule (debug aid)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.dmv_sweep import synth  # noqa: E402
from vlgae_b200 import ops  # noqa: E402
from vlgae_b200._lib import check, lib  # noqa: E402

dev = torch.device("cuda:0")
check(lib().vlgae_dmv_set_schedule(2), "schedule")
# default shapes, or "B,n,len" triples on the command line (all sentences of length len inside tensors padded to n words)
cases = [(1, 40, None, None), (128, 40, "cfg2", None), (4096, 40, None, None), (512, 32, None, None), (512, 20, None, None)]
if len(sys.argv) > 1:
    cases = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
    cases = [(B, n, None, ln) for B, n, ln in cases]
for B, n, ragged, fixed_len in cases:
    md, ma, L = synth(B, n, 7, ragged)
    if fixed_len is not None:
        L[:] = fixed_len
    tmd, tma, tL = [torch.from_numpy(x).to(dev) for x in (md, ma, L)]
    out = ops.ParseBuffers(B, n + 1, dev)
    buf = torch.zeros(16, dtype=torch.int64, device=dev)
    check(lib().vlgae_dmv_set_profile_buffer(buf.data_ptr()), "prof")
    for _ in range(3):
        ops.dmv_parse(tmd, tma, tL, out=out, prepared=True)
    torch.cuda.synchronize()
    buf.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.dmv_parse(tmd, tma, tL, out=out, prepared=True)
    e1.record()
    torch.cuda.synchronize()
    c = buf.cpu().numpy()
    print(f"B={B} n={n} len0={L[0]}: launch {e0.elapsed_time(e1) * 1e3:.1f} us | log: staged {c[0]} inside {c[1] - c[0]} "
          f"outside {c[2] - c[1]} outputs {c[3] - c[2]} total {c[3]} | max: staged {c[4]} chart {c[5] - c[4]} "
          f"backtrace {c[6] - c[5]} total {c[6]}  (cycles)  redo sentences {c[7]}", flush=True)
    if c[8:].any():
        print("   inside width steps, thread 0: before-loop/other %d, loop %d, loads+shuffles %d, finalise %d, to-barrier %d, barrier %d" % tuple(c[8:14]))
check(lib().vlgae_dmv_set_profile_buffer(None), "prof")
check(lib().vlgae_dmv_set_schedule(0), "schedule")
