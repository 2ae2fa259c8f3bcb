#!/usr/bin/env python
"""tools/align_phases.py -- where the MMA-issuing thread of the alignment kernel spends its clocks (debug counters)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vlgae_b200 import _lib  # noqa: E402
from vlgae_b200.alignment import gather_logit_simple  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
A = B = 128
Q, V, D = 82, 1369, 128
vis = torch.randn(A, V, D, generator=g, device=dev)
txt = torch.randn(B, Q, D, generator=g, device=dev)
vm = torch.rand(A, V, generator=g, device=dev) > 0.1
tm = torch.rand(B, Q, generator=g, device=dev) > 0.1
prof = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
lib = _lib.lib()
lib.vlgae_dmv_set_profile_buffer(ctypes.c_void_p(prof.data_ptr()))
for _ in range(3):
    out = gather_logit_simple(vis, vm, txt, tm, split=3, named=False)
torch.cuda.synchronize()
lib.vlgae_dmv_set_profile_buffer(ctypes.c_void_p(0))
pr = prof.view(148, 8).double().cpu()
tot = pr[:, 0].mean().item()
print(f"MMA thread: total {tot:.0f} clk; waiting for ring slots {100 * pr[:, 1].mean().item() / tot:.1f} %, "
      f"for accumulators {100 * pr[:, 2].mean().item() / tot:.1f} %, image tiles {100 * pr[:, 3].mean().item() / tot:.1f} %; "
      f"max/min total {pr[:, 0].max().item():.0f}/{pr[:, 0].min().item():.0f}")
