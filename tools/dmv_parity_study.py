#!/usr/bin/env python
"""tools/dmv_parity_study.py -- three-way error of the log-semiring outputs on the reference's golden vectors.

    python tools/dmv_parity_study.py [--cases dmv_cfg2_full,dmv_len40,...] [--schedules frontier,gather]

For every fixture (produced by the UNMODIFIED reference, tests/golden/gen_golden.py) prints
    |gpu - reference|, |gpu - fp64|, |reference - fp64|
for the arc marginals (d log Z / d attach) and the decision counts (d log Z / d dec), where fp64 is the oracle's
double-precision sweep of the same recurrences, plus log Z relative error and the Viterbi bit-exactness.
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from vlgae_b200 import ops  # noqa: E402
from vlgae_b200._lib import check, lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="dmv_cfg1,dmv_len40,dmv_cfg2_full,dmv_n64,dmv_n128")
    ap.add_argument("--schedules", default="frontier,gather")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    for name in args.cases.split(","):
        g = dict(np.load(os.path.join(ROOT, "tests", "golden", name + ".npz")))
        md, ma = oracle.merge(g["dec"], g["attach"], g["root"])
        L = g["lengths"].astype(np.int64)
        _, gdec64, gatt64 = oracle.dmv_log(md, ma, L, trim=True, f64=True)
        ref_a, ref_d = g["grad_attach"], g["grad_dec"]
        print(f"{name}: B={len(L)} N={md.shape[1]}  |ref-f64| attach {np.abs(ref_a - gatt64).max():.2e} dec {np.abs(ref_d - gdec64).max():.2e}")
        for sched in args.schedules.split(","):
            check(lib().vlgae_dmv_set_schedule({"auto": 0, "frontier": 1, "gather": 2, "role": 3}[sched]), "schedule")
            out = ops.dmv_parse(torch.from_numpy(md).to(dev), torch.from_numpy(ma).to(dev), torch.from_numpy(L).to(dev))
            torch.cuda.synchronize()
            ga, gd = out.gattach.cpu().numpy(), out.gdec.cpu().numpy()
            Z = out.Z.cpu().numpy()
            ea = np.abs(ga - ref_a).reshape(len(L), -1).max(1)
            print(f"   {sched:8s} attach |gpu-ref| {ea.max():.2e} ({(ea > 1e-5).sum()} sentences > 1e-5)  |gpu-f64| {np.abs(ga - gatt64).max():.2e}"
                  f"   dec |gpu-ref| {np.abs(gd - ref_d).max():.2e} |gpu-f64| {np.abs(gd - gdec64).max():.2e}"
                  f"   Z rel {np.abs((Z - g['partition'][:, 0]) / g['partition'][:, 0]).max():.1e}"
                  f"   heads {np.array_equal(out.heads.cpu().numpy(), g['heads'])} best {np.array_equal(out.best.cpu().numpy(), g['max'][:, 0])}")
    check(lib().vlgae_dmv_set_schedule(0), "schedule")


if __name__ == "__main__":
    main()
