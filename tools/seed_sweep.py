#!/usr/bin/env python
"""tools/seed_sweep.py -- the cfg2 launch on the batches the ranks of an N-GPU bench run hold (same length profile, scores of
seed 2 + 1000 r): device time per launch, warm inputs.  Shows how much of the N > 1 headline loss is the scores of a rank."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from vlgae_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
_, _, L0, _ = bench.load_cfg2()
B = len(L0)
for r in range(8):
    if r == 0:
        md, ma, _, _ = bench.load_cfg2()
    else:
        md, ma, _ = bench.make_batch_cpu(B, 2 + 1000 * r)
    t = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (md, ma, L0)]
    out = ops.ParseBuffers(B, md.shape[1], dev)
    for _ in range(20):
        ops.dmv_parse(*t, out=out, prepared=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(500):
        ops.dmv_parse(*t, out=out, prepared=True)
    e1.record()
    torch.cuda.synchronize()
    print(f"rank {r}: {e0.elapsed_time(e1) / 500 * 1e3:.2f} us per launch (warm)")
