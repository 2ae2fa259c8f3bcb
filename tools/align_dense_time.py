"""tools/align_dense_time.py -- time the alignment logits on the dense and on the padded layout (cfg2 shape)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vlgae_b200.alignment import gather_logit_simple
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(99)
A = B = 128
Q, V, D = 82, 1369, 128
vis = torch.randn(A, V, D, generator=g, device=dev)
txt = torch.randn(B, Q, D, generator=g, device=dev)
vm = torch.rand(A, V, generator=g, device=dev) > 0.1
tm = torch.rand(B, Q, generator=g, device=dev) > 0.1
for pad in (False, True):
    for _ in range(3):
        out = gather_logit_simple(vis, vm, txt, tm, named=False, pad_rows=pad)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = gather_logit_simple(vis, vm, txt, tm, named=False, pad_rows=pad)
    e1.record()
    torch.cuda.synchronize()
    print("pad_rows", pad, "ms per call (pack + kernel + allocation)", e0.elapsed_time(e1) / 10, flush=True)
    del out
