import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.dmv_sweep import synth
from vlgae_b200 import ops
from vlgae_b200._lib import check, lib
dev = torch.device("cuda:0")
md, ma, L = synth(1, 40, 7, None)
tmd, tma, tL = [torch.from_numpy(x).to(dev) for x in (md, ma, L)]
out = ops.ParseBuffers(1, 41, dev)
buf = torch.zeros(40, dtype=torch.int64, device=dev)
check(lib().vlgae_dmv_set_profile_buffer(buf.data_ptr()), "prof")
for _ in range(3):
    ops.dmv_parse(tmd, tma, tL, out=out, prepared=True)
torch.cuda.synchronize()
c = buf.cpu().numpy()
names = ["(loop overhead)", "geo+range", "term loop", "combine", "epilogue", "named bar", "(ret)", "syncthreads"]
for r, nm in enumerate(["X ", "CL", "CR"]):
    print("role", nm, {k: int(v) for k, v in zip(names, c[8 + 8 * r: 16 + 8 * r])}, "sum", int(c[8 + 8 * r: 16 + 8 * r].sum()))
print("phase totals", c[:7].tolist())
