import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.dmv_sweep import synth
from vlgae_b200 import ops
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
md, ma, L = synth(B, n, 7, None)
tmd, tma, tL = [torch.from_numpy(x).to(dev) for x in (md, ma, L)]
for _ in range(3):
    Z, gd, ga = ops.dmv_inside_outside(tmd, tma, tL)
torch.cuda.synchronize()
