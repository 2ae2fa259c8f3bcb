"""Small dmv_parse calls (both semirings) for compute-sanitizer runs: python tools/racecheck_parse.py n B"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.dmv_sweep import synth
from vlgae_b200 import ops
dev = torch.device("cuda:0")
n, B = int(sys.argv[1]), int(sys.argv[2])
md, ma, L = synth(B, n, 7, "cfg2" if n >= 8 else None)
out = ops.dmv_parse(*[torch.from_numpy(x).to(dev) for x in (md, ma, L)], want_arcs=True)
torch.cuda.synchronize()
print("ok", float(out.Z[0]), int(out.heads.sum()))
