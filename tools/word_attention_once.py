"""tools/word_attention_once.py -- a few launches of the word -> factor attention at the cfg2 shape (for ncu captures)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vlgae_b200.alignment import word_factor_attention
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
B, V, n, D, H = 128, 1369, 40, 128, 256
vis = (torch.randn(B, V, D, generator=g, device=dev) * 0.2).requires_grad_()
txt = (torch.randn(B, n, D, generator=g, device=dev) * 0.2).requires_grad_()
mid = torch.randn(B, V, H, generator=g, device=dev).requires_grad_()
go = torch.randn(B, n, H, generator=g, device=dev)
for _ in range(3):
    out = word_factor_attention(vis, txt, mid)
    torch.autograd.grad(out, [vis, txt, mid], go)
torch.cuda.synchronize()
