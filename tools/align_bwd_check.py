#!/usr/bin/env python
"""tools/align_bwd_check.py -- vlgae_align_logits_backward against torch's fp32 contractions (error + timing)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vlgae_b200._lib import check, lib  # noqa: E402

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
shapes = [(2, 150, 3, 20, 128), (3, 300, 4, 82, 128), (2, 129, 2, 33, 64), (2, 40, 2, 5, 8)]
if len(sys.argv) > 1 and sys.argv[1] == "big":
    shapes = [(128, 1369, 128, 82, 128)]
for A, V, B, Q, D in shapes:
    g_ = torch.Generator(device=dev).manual_seed(A + V)
    vis = torch.randn(A, V, D, generator=g_, device=dev)
    txt = torch.randn(B, Q, D, generator=g_, device=dev)
    vm = torch.rand(A, V, generator=g_, device=dev) > 0.2
    tm = torch.rand(B, Q, generator=g_, device=dev) > 0.2
    g = torch.randn(B, A, Q, V, generator=g_, device=dev)
    gv = torch.full((A, V, D), 7.0, device=dev)
    gt = torch.full((B, Q, D), 7.0, device=dev)
    need = lib().vlgae_align_workspace_bytes(A, V, B, Q, D)
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    vmu, tmu = vm.view(torch.uint8), tm.view(torch.uint8)

    def run():
        check(lib().vlgae_align_logits_backward(g.data_ptr(), V, vis.data_ptr(), vmu.data_ptr(), txt.data_ptr(), tmu.data_ptr(),
                                                A, V, B, Q, D, 3, gv.data_ptr(), gt.data_ptr(), ws.data_ptr(), ws.numel(),
                                                torch.cuda.current_stream().cuda_stream), "bwd")
    run()
    torch.cuda.synchronize()
    if A * B * Q * V < 5e7:
        gm = g * vm[None, :, None, :] * tm[:, None, :, None]
        rv = torch.einsum("baqv,bqd->avd", gm, txt)
        rt = torch.einsum("baqv,avd->bqd", gm, vis)
        print(f"A={A} V={V} B={B} Q={Q} D={D}: |d_vis err| {float((gv - rv).abs().max()):.3e} (scale {float(rv.abs().max()):.1f})  "
              f"|d_txt err| {float((gt - rt).abs().max()):.3e} (scale {float(rt.abs().max()):.1f})")
    else:
        def timed(pv, pt):
            def f():
                check(lib().vlgae_align_logits_backward(g.data_ptr(), V, vis.data_ptr(), vmu.data_ptr(), txt.data_ptr(),
                                                        tmu.data_ptr(), A, V, B, Q, D, int(os.environ.get("SPLIT", "3")), pv, pt, ws.data_ptr(), ws.numel(),
                                                        torch.cuda.current_stream().cuda_stream), "bwd")
            f()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                f()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / 3
        import ctypes
        prof = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
        lib().vlgae_dmv_set_profile_buffer(ctypes.c_void_p(prof.data_ptr()))
        for name, pv, pt in (("d_txt", None, gt.data_ptr()), ("d_vis", gv.data_ptr(), None)):
            prof.zero_()
            timed(pv, pt)
            pr = prof.view(148, 8).double().cpu()
            pr = pr[pr[:, 0] > 0]
            tot = pr[:, 0].mean().item()
            print(f"  {name}: MMA warp {tot:.0f} clk per launch; waiting for the operand tile {100 * pr[:, 1].mean().item() / tot:.0f} %, "
                  f"for the converted g image {100 * pr[:, 2].mean().item() / tot:.0f} %")
        lib().vlgae_dmv_set_profile_buffer(ctypes.c_void_p(0))
        print(f"A={A} V={V} B={B} Q={Q} D={D}: d_txt {timed(None, gt.data_ptr()):.3f} ms, d_vis {timed(gv.data_ptr(), None):.3f} ms, "
              f"both {timed(gv.data_ptr(), gt.data_ptr()):.3f} ms  (gradient: {g.numel() * 4 / 2**30:.2f} GiB)")
