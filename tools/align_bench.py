#!/usr/bin/env python
"""tools/align_bench.py -- time the alignment kernel at the cfg2 shape (A = B = 128, Q = 82, V = 1369, D = 128)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vlgae_b200.alignment import gather_logit_reduced, gather_logit_simple, max_over_factors  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--A", type=int, default=128)
ap.add_argument("--B", type=int, default=128)
ap.add_argument("--Q", type=int, default=82)
ap.add_argument("--V", type=int, default=1369)
ap.add_argument("--D", type=int, default=128)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--mask", choices=("random", "prefix"), default="random",
                help="random: 10 %% of queries masked at random; prefix: per-caption lengths (ROOT + padding masked), as in training")
ap.add_argument("--quick", action="store_true", help="only the split-3 padded configuration, no torch comparison")
args = ap.parse_args()
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
vis = torch.randn(args.A, args.V, args.D, generator=g, device=dev)
txt = torch.randn(args.B, args.Q, args.D, generator=g, device=dev)
vm = torch.rand(args.A, args.V, generator=g, device=dev) > 0.1
tm = torch.rand(args.B, args.Q, generator=g, device=dev) > 0.1
if args.mask == "prefix":
    half = args.Q // 2
    ln = torch.randint(4, half, (args.B, 1), generator=g, device=dev)
    pos = torch.arange(half, device=dev)[None]
    keep = (pos >= 1) & (pos <= ln)
    tm = torch.cat([keep, keep], 1) if args.Q == 2 * half else torch.cat([keep, keep, keep[:, :1]], 1)
out_bytes = args.A * args.B * args.Q * args.V * 4
flops = 2.0 * args.A * args.B * args.Q * args.V * args.D
for split, pad in (((3, True),) if args.quick else ((3, True), (3, False), (1, True))):
    for _ in range(2):
        out = gather_logit_simple(vis, vm, txt, tm, split=split, named=False, pad_rows=pad)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        out = gather_logit_simple(vis, vm, txt, tm, split=split, named=False, pad_rows=pad)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    print(f"split={split} pad_rows={pad}: {ms:.3f} ms  {out_bytes / ms / 1e6:.0f} GB/s written  {flops * (3 if split == 3 else 1) / ms / 1e9:.0f} "
          f"TFLOP/s issued (bf16)  [{args.B}x{args.A}x{args.Q}x{args.V}, out {out_bytes / 2**30:.2f} GiB]")
def timed(fn, iters=args.iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


marg = torch.rand(args.B, args.Q, generator=g, device=dev) + 0.1
ms_fused = timed(lambda: max_over_factors(vis, vm, txt, tm))
ms_mat = timed(lambda: gather_logit_simple(vis, vm, txt, tm, named=False).max(dim=-1).values)
print(f"max over V, fused epilogue (no [B,A,Q,V] tensor): {ms_fused:.3f} ms  {flops * 3 / ms_fused / 1e9:.0f} TFLOP/s issued;  "
      f"materialise + torch max: {ms_mat:.3f} ms")
print(f"gather_logit_reduced end to end: {timed(lambda: gather_logit_reduced(vis, vm, txt, tm, marg)):.3f} ms")
if args.quick:
    sys.exit(0)
# reference arithmetic for context: fp32 einsum + 2 masked fills (what joint.py:413-418 runs on the GPU)
torch.backends.cuda.matmul.allow_tf32 = False
for _ in range(2):
    ref = torch.einsum("avd,bqd->baqv", vis, txt)
    ref.masked_fill_(~vm[None, :, None, :], -1e20)
    ref.masked_fill_(~tm[:, None, :, None], -1e20)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ref = torch.einsum("avd,bqd->baqv", vis, txt)
ref.masked_fill_(~vm[None, :, None, :], -1e20)
ref.masked_fill_(~tm[:, None, :, None], -1e20)
e1.record()
torch.cuda.synchronize()
print(f"torch einsum fp32 + 2 masked_fill_ (the reference's GPU path): {e0.elapsed_time(e1):.3f} ms")
err = (out - ref)[ref != -1e20].abs().max().item()
print(f"max |ours(split=1) - torch fp32| = {err:.3e}")
out3 = gather_logit_simple(vis, vm, txt, tm, split=3, named=False)
print(f"max |ours(split=3) - torch fp32| = {(out3 - ref)[ref != -1e20].abs().max().item():.3e};  mask pattern equal: "
      f"{bool(((out3 == -1e20) == (ref == -1e20)).all())}")
