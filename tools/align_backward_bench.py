#!/usr/bin/env python
"""tools/align_backward_bench.py -- forward + backward of gather_logit_simple at the cfg2 shape, ours vs the reference formula."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vlgae_b200.alignment import gather_logit_simple  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
A = B = 128
Q, V, D = 82, 1369, 128
vis = torch.randn(A, V, D, generator=g, device=dev).requires_grad_()
txt = torch.randn(B, Q, D, generator=g, device=dev).requires_grad_()
vm = torch.rand(A, V, generator=g, device=dev) > 0.1
tm = torch.rand(B, Q, generator=g, device=dev) > 0.1
torch.backends.cuda.matmul.allow_tf32 = False


def ours():
    out = gather_logit_simple(vis, vm, txt, tm, named=False)
    out.backward(torch.ones_like(out))


def ref():
    out = torch.einsum("avd,bqd->baqv", vis, txt)
    out = out.masked_fill(~vm[None, :, None, :], -1e20).masked_fill(~tm[:, None, :, None], -1e20)
    out.backward(torch.ones_like(out))


for name, fn in (("vlgae_b200", ours), ("reference formula (torch)", ref)):
    fn()
    torch.cuda.synchronize()
    vis.grad = txt.grad = None
    t0 = time.perf_counter()
    for _ in range(3):
        fn()
        vis.grad = txt.grad = None
    torch.cuda.synchronize()
    print(f"{name}: forward + backward {1e3 * (time.perf_counter() - t0) / 3:.1f} ms, peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
    torch.cuda.reset_peak_memory_stats()
