#!/usr/bin/env python
"""tools/dmv_sweep.py -- time vlgae_dmv_parse over (batch, length, launch tuning) on one GPU.

    python tools/dmv_sweep.py [--configs cfg2,bulk,cfg3] [--tunings auto,1x64,4x128,...] [--errors]

Prints one line per (config, tuning): us per launch, sentences/s, fraction of the measured MUFU peak.
With --errors also prints max |marginal - f64 oracle| and max |marginal - f32 oracle| (noise floor study).
"""
import argparse
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from vlgae_b200 import ops  # noqa: E402
from vlgae_b200._lib import check, lib  # noqa: E402


def synth(B, n, seed, ragged):
    g = torch.Generator().manual_seed(seed)
    dec = torch.randn(B, n, 2, 2, 2, generator=g).log_softmax(-1)
    attach = torch.randn(B, n, n, 2, generator=g).log_softmax(2)
    root = torch.randn(B, n, generator=g).log_softmax(-1)
    if ragged == "cfg2":
        L = torch.randint(4, n + 1, (B,), generator=g).sort(descending=True).values
        L[0] = n
    elif ragged == "coco":  # cfg5: clamp(round(N(11, 4^2)), 3, 40), processed length-sorted
        L = torch.clamp((torch.randn(B, generator=g) * 4 + 11).round(), 3, n).long().sort(descending=True).values
    else:
        L = torch.full((B,), n, dtype=torch.long)
    md, ma = oracle.merge(dec.numpy(), attach.numpy(), root.numpy())
    return md, ma, L.numpy().astype(np.int64)


CONFIGS = {
    "cfg2": (128, 40, "cfg2"), "cfg2x8": (1024, 40, "cfg2"), "bulk": (16384, 40, "coco"), "bulk100k": (100000, 40, "coco"),
    "bulk40": (8192, 40, "cfg2"), "n8": (512, 8, None), "n16": (512, 16, None), "n32": (512, 32, None),
    "n20": (512, 20, None), "n24": (512, 24, None), "n28": (512, 28, None), "n36": (512, 36, None), "n40": (512, 40, None),
    "n64": (512, 64, None), "n128": (512, 128, None), "full40": (4096, 40, None),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="cfg2,bulk")
    ap.add_argument("--tunings", default="auto")
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--errors", action="store_true")
    ap.add_argument("--schedules", default="auto", help="comma list of auto,frontier,gather,role")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    L_ = lib()
    ms, nops = ctypes.c_float(), ctypes.c_double()
    peak = 0.0
    for _ in range(3):
        check(L_.vlgae_microbench_mufu(4000, ctypes.byref(ms), ctypes.byref(nops), None), "mufu")
        peak = max(peak, nops.value / (ms.value * 1e-3))
    print(f"MUFU peak {peak / 1e9:.0f} Gop/s")
    for cname in args.configs.split(","):
        B, n, ragged = CONFIGS[cname]
        md, ma, L = synth(B, n, 7, ragged)
        N = L.astype(np.float64) + 1
        mufu = float((2 * (N ** 3 - N) + 3 * N * (N - 1)).sum())
        tmd, tma, tL = torch.from_numpy(md).to(dev), torch.from_numpy(ma).to(dev), torch.from_numpy(L).to(dev)
        out = ops.ParseBuffers(B, n + 1, dev)
        if args.errors:
            nb = min(B, 64)
            _, _, g64 = oracle.dmv_log(md[:nb], ma[:nb], L[:nb], f64=True, trim=True)
            _, _, g32 = oracle.dmv_log(md[:nb], ma[:nb], L[:nb], trim=True)
        for sched, tn in [(s_, t_) for s_ in args.schedules.split(",") for t_ in args.tunings.split(",")]:
            check(L_.vlgae_dmv_set_schedule({"auto": 0, "frontier": 1, "gather": 2, "role": 3}[sched]), "schedule")
            for _ in range(3):
                ops.dmv_parse(tmd, tma, tL, out=out, prepared=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                ops.dmv_parse(tmd, tma, tL, out=out, prepared=True)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / args.iters
            line = (f"{cname:8s} B={B:6d} n={n:3d} sched={sched:8s} tuning={tn:9s} {us:10.1f} us/launch  {B / us:8.3f} Msent/s  "
                    f"mufu_frac={mufu / (us * 1e-6) / peak:.3f}")
            if args.errors:
                gpu = out.gattach[:nb].cpu().numpy()
                line += f"  |gpu-f64|={np.abs(gpu - g64).max():.2e} |gpu-f32|={np.abs(gpu - g32).max():.2e} |f32-f64|={np.abs(g32 - g64).max():.2e}"
            print(line, flush=True)
    check(L_.vlgae_dmv_set_schedule(0), "schedule")


if __name__ == "__main__":
    main()
