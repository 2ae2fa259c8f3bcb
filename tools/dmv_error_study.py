"""Where do the marginals of the CUDA path differ from the reference's golden vectors / the fp64 oracle?"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
from vlgae_b200 import ops
dev = torch.device("cuda:0")
for name in ["dmv_len40", "dmv_cfg1"]:
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", name + ".npz")))
    md, ma, L = g["merged_dec"], g["merged_attach"], g["lengths"]
    Z, gd, ga = ops.dmv_inside_outside(torch.from_numpy(md).to(dev), torch.from_numpy(ma).to(dev), torch.from_numpy(L).to(dev))
    ga = ga.cpu().numpy(); Z = Z.cpu().numpy()
    Z64, _, ga64 = oracle.dmv_log(md, ma, L, f64=True, trim=True)
    Z32, _, ga32 = oracle.dmv_log(md, ma, L, trim=True)
    ref = g["grad_attach"]
    print(name, "Z gpu-ref", np.abs(Z - g["partition"][:, 0]).max(), "Z gpu-f64", np.abs(Z - Z64).max(), "Z ref-f64", np.abs(g["partition"][:, 0] - Z64).max())
    for b in range(min(len(L), 4)):
        d_ref = np.abs(ga[b] - ref[b]); d64 = np.abs(ga[b] - ga64[b]); r64 = np.abs(ref[b] - ga64[b]); o_ref = np.abs(ga32[b] - ref[b])
        k = np.unravel_index(d_ref.argmax(), d_ref.shape)
        print(f"  b={b} len={L[b]} |gpu-ref| max {d_ref.max():.2e} (#>5e-6: {(d_ref > 5e-6).sum()}) at {k} val {ref[b][k]:.4f} | |gpu-f64| {d64.max():.2e} |ref-f64| {r64.max():.2e} |oracle32-ref| {o_ref.max():.2e}"
              f" | rel at max {d_ref.max() / max(ref[b][k], 1e-30):.2e}")
