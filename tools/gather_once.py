"""tools/gather_once.py -- a few gather-schedule launches of one shape (for ncu captures).
    python tools/gather_once.py [B] [n] [log|max|both]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.dmv_sweep import synth
from vlgae_b200 import ops
from vlgae_b200._lib import check, lib
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
what = sys.argv[3] if len(sys.argv) > 3 else "both"
check(lib().vlgae_dmv_set_schedule(2), "schedule")
md, ma, L = synth(B, n, 7, None)
tmd, tma, tL = [torch.from_numpy(x).to(dev) for x in (md, ma, L)]
out = ops.ParseBuffers(B, n + 1, dev)
for _ in range(3):
    if what == "log":
        ops.dmv_inside_outside(tmd, tma, tL)
    elif what == "max":
        ops.dmv_viterbi(tmd, tma, tL)
    else:
        ops.dmv_parse(tmd, tma, tL, out=out, prepared=True)
torch.cuda.synchronize()
