#!/usr/bin/env python
"""tools/deptree_bench.py -- time DependencyCRF (MBR decoding path, ldndmv.py:294-299) at the cfg2 shape."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vlgae_b200 import deptree  # noqa: E402
from vlgae_b200.torch_struct.semirings import LogSemiring, MaxSemiring  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
for B, n in ((128, 40), (1024, 40), (512, 16)):
    N = n + 1
    arc = torch.rand(B, N, N, generator=g, device=dev)
    L = torch.randint(4, n + 1, (B,), generator=g, device=dev).sort(descending=True).values
    L[0] = n
    for sem, S in (("log", LogSemiring), ("max", MaxSemiring)):
        for _ in range(3):
            out = deptree.run(arc, L, S, True)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            out = deptree.run(arc, L, S, True)
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) / 20 * 1e3
        print(f"DependencyCRF {sem:3s} B={B:5d} n={n:3d}: {us:9.1f} us/call  {B / us:.3f} Msent/s")
