"""tools/frontier_phases.py -- per-warp clocks of the frontier kernel's reverse sweep (build with -DVLGAE_PROF_DETAIL)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.dmv_sweep import synth
from vlgae_b200 import ops
from vlgae_b200._lib import check, lib
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
md, ma, L = synth(1, n, 7, None)
tmd, tma, tL = [torch.from_numpy(x).to(dev) for x in (md, ma, L)]
buf = torch.zeros(8 + 32 * 4, dtype=torch.int64, device=dev)
check(lib().vlgae_dmv_set_profile_buffer(buf.data_ptr()), "prof")
for _ in range(3):
    ops.dmv_inside_outside(tmd, tma, tL)
torch.cuda.synchronize()
c = buf.cpu().numpy()
print("phase totals (cycles): staged %d inside %d outside %d outputs %d" % (c[0], c[1] - c[0], c[2] - c[1], c[3] - c[2]))
print("warp:  A' tasks | A' barrier | B' tasks | B' barrier")
for w in range(16):
    print(w, c[8 + 4 * w: 12 + 4 * w].tolist())
