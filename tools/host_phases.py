#!/usr/bin/env python
"""tools/host_phases.py -- phase timers of the sentence-0 CTAs inside vlgae_dmv_parse_host (zero-copy path, cfg2 batch).
VLGAE_PROF_ALL=1 adds the start / end time of every work item; that needs a library built with -DVLGAE_TIMELINE
(NVCC_EXTRA=-DVLGAE_TIMELINE python __graft_entry__.py --force)."""
import ctypes
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from vlgae_b200._lib import check, lib  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
md0, ma0, L0, _ = bench.load_cfg2()
B, N = md0.shape[0], md0.shape[1]
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()  # noqa: E731
h_md, h_ma, h_L = pin(md0), pin(ma0), pin(L0)
h_Z, h_best = torch.empty(B).pin_memory(), torch.empty(B).pin_memory()
h_gatt = torch.empty((B, N, N, 2)).pin_memory()
h_gdec = torch.empty((B, N, 2, 2, 2)).pin_memory()
h_heads = torch.empty((B, N), dtype=torch.int64).pin_memory()
L_ = lib()
buf = torch.zeros(8 + 4 * B, dtype=torch.int64, device=dev)
stream = torch.cuda.current_stream().cuda_stream


def call():
    nograd = bool(os.environ.get("NO_GRAD_OUT"))  # experiment: no marginals written back (the sweeps still run if asked)
    check(L_.vlgae_dmv_parse_host(h_md.data_ptr(), h_ma.data_ptr(), h_L.data_ptr(), B, N, ctypes.c_float(-1e12),
                                  h_Z.data_ptr(), None if nograd else h_gdec.data_ptr(), None if nograd else h_gatt.data_ptr(),
                                  h_best.data_ptr(), h_heads.data_ptr(), stream), "parse_host")


for _ in range(5):
    call()
t0 = time.perf_counter()
for _ in range(200):
    call()
print(f"mean of 200 calls: {(time.perf_counter() - t0) / 200 * 1e6:.1f} us")
check(L_.vlgae_dmv_set_profile_buffer(buf.data_ptr()), "prof")
for _ in range(3):
    t0 = time.perf_counter()
    call()
    dt = (time.perf_counter() - t0) * 1e6
    c = buf.cpu().numpy()
    print(f"call {dt:.1f} us | log: staged {c[0]} inside {c[1] - c[0]} outside {c[2] - c[1]} outputs {c[3] - c[2]} total {c[3]} | "
          f"max: staged {c[4]} chart {c[5] - c[4]} backtrace {c[6] - c[5]} total {c[6]} (cycles)")
if os.environ.get("VLGAE_PROF_ALL"):
    tl = buf.cpu().numpy()[8:].reshape(2, B, 2).astype(np.float64)
    t0 = tl[:, :, 0].min()
    tl = (tl - t0) / 1e3
    print("sentence  len | log start  end | max start  end   (us after the first CTA started)")
    for b in list(range(0, B, 8)) + [B - 1]:
        print(f"{b:8d} {int(L0[b]):4d} | {tl[0, b, 0]:8.1f} {tl[0, b, 1]:6.1f} | {tl[1, b, 0]:8.1f} {tl[1, b, 1]:6.1f}")
    print("last end: log %.1f (sentence %d), max %.1f (sentence %d)" % (tl[0, :, 1].max(), tl[0, :, 1].argmax(), tl[1, :, 1].max(), tl[1, :, 1].argmax()))
check(L_.vlgae_dmv_set_profile_buffer(None), "prof")
