#!/usr/bin/env python
"""bench.py -- sentences/sec of the DMV hot path (inside + outside + Viterbi, len <= 40) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one cfg2 batch per GPU (BASELINE.json configs[1]: 128 captions,
ragged lengths 4..40 sorted descending, fp32 merged score tensors): log-semiring inside, the explicit outside
sweep (arc + decision expected counts) and max-semiring Viterbi with head decode, all in ONE kernel launch
(vlgae_dmv_parse).  Sentences shard across GPUs with no data-path collective (weak scaling: 128 per GPU).

Prints ONE JSON line on rank 0 (see README / DESIGN.md for the keys).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

BATCH_PER_GPU = 128
MAX_LEN = 40
MASK_ZERO = -1e12
L2_BYTES = 126 * 1024 * 1024


# ----------------------------------------------------------------------------------------------------------
# workload
# ----------------------------------------------------------------------------------------------------------
def make_lengths(B, seed):
    import torch

    g = torch.Generator().manual_seed(seed)
    L = torch.randint(4, MAX_LEN + 1, (B,), generator=g).sort(descending=True).values
    L[0] = MAX_LEN
    return L


def make_batch_cpu(B, seed):
    """Merged score tensors of one batch, built on the host with the oracle's merge."""
    import torch

    import oracle

    g = torch.Generator().manual_seed(seed)
    n = MAX_LEN
    dec = torch.randn(B, n, 2, 2, 2, generator=g).log_softmax(-1)
    attach = torch.randn(B, n, n, 2, generator=g).log_softmax(2)
    root = torch.randn(B, n, generator=g).log_softmax(-1)
    md, ma = oracle.merge(dec.numpy(), attach.numpy(), root.numpy())
    return md, ma, make_lengths(B, seed).numpy().astype(np.int64)


def work_counts(lengths):
    """Algorithmic work of inside + outside + Viterbi: split-point terms T(N) = N^3 - N per
    chart sweep; MUFU ops = 2 T + 3 N (N - 1); FP32-pipe ops ~ 10 T; HBM bytes = 16 N^2 + 64 N + 8 len + 16."""
    N = np.asarray(lengths, dtype=np.float64) + 1
    T = N ** 3 - N
    return dict(mufu=float((2 * T + 3 * N * (N - 1)).sum()), fp32=float((10 * T).sum()),
                hbm_bytes=float((16 * N * N + 64 * N + 8 * (N - 1) + 16).sum()))


# ----------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.timed = False
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                mhz = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
                try:
                    r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((self.timed, mhz))
                if self.timed:
                    for bit, name in self.REASONS.items():
                        if r & bit:
                            self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()

    def summary(self):
        timed = [m for t, m in self.samples if t] or [m for _, m in self.samples[-3:]]
        return {"sm_mhz": float(np.median(timed)) if timed else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len([1 for t, _ in self.samples if t])}


# ----------------------------------------------------------------------------------------------------------
# CPU legs (oracle = C port of the reference's algorithm; the Python reference cannot travel to the GPU box)
# ----------------------------------------------------------------------------------------------------------
def oracle_step(md, ma, L, threads):
    """inside + outside + Viterbi with the oracle, sentences sharded over `threads` host threads
    (ctypes releases the GIL, so the C sweeps run concurrently)."""
    import oracle

    B = len(L)
    if threads <= 1 or B < 2:
        oracle.dmv_log(md, ma, L, trim=True)
        oracle.dmv_viterbi(md, ma, L, trim=True)
        return
    from concurrent.futures import ThreadPoolExecutor

    # interleave so every shard gets the same mix of lengths
    shards = [np.arange(k, B, threads) for k in range(min(threads, B))]

    def run(idx):
        a, b, c = np.ascontiguousarray(md[idx]), np.ascontiguousarray(ma[idx]), np.ascontiguousarray(L[idx])
        oracle.dmv_log(a, b, c, trim=True)
        oracle.dmv_viterbi(a, b, c, trim=True)

    with ThreadPoolExecutor(len(shards)) as ex:
        list(ex.map(run, shards))


def cpu_baseline(md, ma, L, budget_s=10.0):
    """Scalar oracle port on ONE host core over the whole cfg2 batch, repeated for ~budget_s seconds."""
    import oracle

    oracle.lib()
    t0 = time.perf_counter()
    reps = 0
    while True:
        oracle_step(md, ma, L, 1)
        reps += 1
        el = time.perf_counter() - t0
        if el >= budget_s:
            break
    return {"value": len(L) * reps / el, "unit": "sentences/s", "cores": 1, "kind": "port",
            "sample": f"{reps} x the full cfg2 batch ({len(L)} sentences) in {el:.1f} s, oracle/dmv_oracle.c "
                      f"(inside+outside+Viterbi, fp32)"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle

    oracle.lib()
    threads = os.cpu_count() or 1
    md, ma, L = make_batch_cpu(BATCH_PER_GPU, 2)
    # calibrate, then bound each step so that the whole run ends within ~2 minutes
    t0 = time.perf_counter()
    oracle_step(md[:threads * 2], ma[:threads * 2], L[:threads * 2], threads)
    rate = (threads * 2) / max(time.perf_counter() - t0, 1e-6)
    budget = 90.0 / max(args.steps + args.warmup, 1)
    S = int(max(min(BATCH_PER_GPU, rate * budget), min(threads, BATCH_PER_GPU)))
    idx = np.linspace(0, BATCH_PER_GPU - 1, S).round().astype(int)  # same length mix as the full batch
    smd, sma, sL = np.ascontiguousarray(md[idx]), np.ascontiguousarray(ma[idx]), np.ascontiguousarray(L[idx])
    for _ in range(args.warmup):
        oracle_step(smd, sma, sL, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_step(smd, sma, sL, threads)
    el = time.perf_counter() - t0
    value = S * args.steps / el
    line = {
        "impl": "reference", "metric": "sentences/sec (inside+outside+Viterbi, len<=40)", "value": value,
        "unit": "sentences/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": el / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2: 128 captions, len 4..40 ragged sorted desc (BASELINE.json configs[1]), "
                               "DMV inside+outside+Viterbi", "sample_sentences_per_step": S},
        "cpu_baseline": {"value": value, "unit": "sentences/s", "cores": threads, "kind": "port",
                         "sample": f"{S} of the 128 cfg2 sentences per step (same length mix), {threads} host threads; "
                                   "the reference is pure Python/PyTorch and cannot travel to the GPU box, so this "
                                   "is oracle/dmv_oracle.c, the C restatement pinned to it"},
        "e2e": {"value": value, "unit": "sentences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------------------
# alignment leg (second kernel of the hot path; reported beside the headline, not part of `value`)
# ----------------------------------------------------------------------------------------------------------
def alignment_leg(dev, iters=10):
    """gather_logit_simple at the cfg2 shape: A = B = 128 images/captions, Q = 2 * 41 queries, V = 36 + 36^2 + 36 + 1
    = 1369 factors, D = 128.  HBM-write-bound: 4 * B * A * Q * V bytes must be written once."""
    import torch

    import oracle
    from vlgae_b200.alignment import gather_logit_simple, max_over_factors

    A = B = BATCH_PER_GPU
    Q, V, D = 2 * (MAX_LEN + 1), 36 + 36 * 36 + 36 + 1, 128
    g = torch.Generator(device=dev).manual_seed(99)
    vis = torch.randn(A, V, D, generator=g, device=dev)
    txt = torch.randn(B, Q, D, generator=g, device=dev)
    vm = torch.rand(A, V, generator=g, device=dev) > 0.1
    tm = torch.rand(B, Q, generator=g, device=dev) > 0.1
    out = None
    for _ in range(2):
        out = gather_logit_simple(vis, vm, txt, tm, named=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = gather_logit_simple(vis, vm, txt, tm, named=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    # parity on a corner block against the oracle (numpy fp32 restatement of joint.py:406-419)
    nb, na = 2, 3
    want = oracle.gather_logit_simple(vis[:na].cpu().numpy(), vm[:na].cpu().numpy(), txt[:nb].cpu().numpy(),
                                      tm[:nb].cpu().numpy())
    got = out[:nb, :na].cpu().numpy()
    masked = want == -1e20
    err = float(np.abs(got - want)[~masked].max())
    peak = None
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        src = "MEASURED_PEAKS.json hbm_gbs (measured)"
    except Exception:  # noqa: BLE001
        peak, src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    out_bytes = 4.0 * B * A * Q * V
    gbs = out_bytes / (ms * 1e-3) / 1e9
    # gather_logit_reduced's first half: max over V fused into the epilogue, no [B,A,Q,V] tensor (tensor-bound form)
    ref_max = out.max(dim=-1).values
    del out
    for _ in range(2):
        maxv, _ = max_over_factors(vis, vm, txt, tm)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        maxv, _ = max_over_factors(vis, vm, txt, tm)
    e1.record()
    torch.cuda.synchronize()
    ms_red = e0.elapsed_time(e1) / iters
    flops = 3 * 2.0 * A * B * Q * V * D
    try:
        tpeak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]
        tsrc = "MEASURED_PEAKS.json bf16_tflops (measured, burst)"
    except Exception:  # noqa: BLE001
        tpeak, tsrc = 1500.0, "fallback 1.5 PFLOP/s dense bf16 (B200_PROFILING.md)"
    # backward of the materialised logits (both transposed contractions, tcgen05): upstream gradient = ones
    from vlgae_b200._lib import check as _check, lib as _lib
    gup = torch.ones((B, A, Q, V), dtype=torch.float32, device=dev)
    gvis, gtxt = torch.empty_like(vis), torch.empty_like(txt)
    ws = torch.empty(_lib().vlgae_align_workspace_bytes(A, V, B, Q, D), dtype=torch.uint8, device=dev)
    vmu, tmu = vm.view(torch.uint8), tm.view(torch.uint8)

    def bwd():
        _check(_lib().vlgae_align_logits_backward(gup.data_ptr(), V, vis.data_ptr(), vmu.data_ptr(), txt.data_ptr(), tmu.data_ptr(),
                                                  A, V, B, Q, D, 3, gvis.data_ptr(), gtxt.data_ptr(), ws.data_ptr(), ws.numel(),
                                                  torch.cuda.current_stream(dev).cuda_stream), "vlgae_align_logits_backward")
    bwd()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        bwd()
    e1.record()
    torch.cuda.synchronize()
    ms_bwd = e0.elapsed_time(e1) / iters
    want_t = ((vis * vm.unsqueeze(-1)).sum((0, 1)).unsqueeze(0).unsqueeze(0) * tm.unsqueeze(-1)).cpu()  # g = 1: sum of kept vis rows
    bwd_err = float((gtxt.cpu() - want_t).abs().max() / want_t.abs().max())
    del gup
    backward = {"workload": "vlgae_align_logits_backward (d vis and d txt, gradient streamed once each), same shape",
                "ms": ms_bwd, "gradient_gb_per_s": 2 * out_bytes / (ms_bwd * 1e-3) / 1e9,
                "rel_err_d_txt_vs_closed_form": bwd_err}
    reduced = {"workload": "vlgae_align_max_over_factors (max over V in the epilogue; joint.py:421-428), same shape",
               "ms": ms_red, "bit_identical_to_max_of_materialised": bool(torch.equal(maxv, ref_max)),
               "roofline": {"bound": "tensor", "achieved": flops / (ms_red * 1e-3) / 1e12, "peak": tpeak, "unit": "TFLOP/s",
                            "frac": flops / (ms_red * 1e-3) / 1e12 / tpeak, "traffic": None, "peak_source": tsrc,
                            "note": "flops = the 3 bf16 MMAs of the hi/lo split per logit"}}
    return {
        "workload": f"gather_logit_simple A={A} V={V} B={B} Q={Q} D={D} (cfg2), bf16 hi/lo split x3 on tcgen05, "
                    "rows padded to 8 floats", "ms": ms, "captions_per_s": B / (ms * 1e-3),
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                     # dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu capture, profiles/r1_align_kernel.txt)
                     "traffic": 7.61e9, "peak_source": src, "algorithmic_bytes_per_launch": out_bytes,
                     "kernel": "align_gemm_kernel (+ align_pack_kernel x2)"},
        "tensor_tflops_issued": 3 * 2.0 * A * B * Q * V * D / (ms * 1e-3) / 1e12,
        "parity": {"mask_pattern_equal": bool(((got == -1e20) == masked).all()), "max_abs_err_vs_fp32_oracle": err},
        "gpu_launches_per_call": 3,
        "reduced": reduced,
        "backward": backward,
    }


# ----------------------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------------------
def run_b200_arm(args):
    import torch

    from vlgae_b200 import ops
    from vlgae_b200._lib import check, lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the b200 arm has no CPU fallback; use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    L_ = lib()
    B, N = BATCH_PER_GPU, MAX_LEN + 1

    # ---- inputs: a pool of distinct batches whose footprint exceeds L2, so every step reads cold data ----
    md0, ma0, L0 = make_batch_cpu(B, 2 + 1000 * rank)
    step_bytes = md0.nbytes + ma0.nbytes + L0.nbytes + md0.nbytes + ma0.nbytes + B * 4 * 2 + B * N * 8
    pool_n = int(np.ceil(2.2 * L2_BYTES / step_bytes))
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    from vlgae_b200.torch_struct import DMV1o

    # same construction as make_batch_cpu, different draws, built on the device for the whole pool at once
    P = (pool_n - 1) * B
    dec = torch.randn(P, MAX_LEN, 2, 2, 2, generator=gen, device=dev).log_softmax(-1)
    att = torch.randn(P, MAX_LEN, MAX_LEN, 2, generator=gen, device=dev).log_softmax(2)
    root = torch.randn(P, MAX_LEN, generator=gen, device=dev).log_softmax(-1)
    pmd, pma = DMV1o.merge(dec, att, root)
    del dec, att, root
    pool = []
    for k in range(pool_n):
        if k == 0:
            md, ma = torch.from_numpy(md0).to(dev), torch.from_numpy(ma0).to(dev)
        else:
            md, ma = pmd[(k - 1) * B:k * B], pma[(k - 1) * B:k * B]
        pool.append((md, ma, torch.from_numpy(L0).to(dev), ops.ParseBuffers(B, N, dev)))
    torch.cuda.synchronize()

    def step(k):
        md, ma, L, out = pool[k % pool_n]
        ops.dmv_parse(md, ma, L, out=out, prepared=True)

    # ---- parity gate on the timed configuration ----
    import oracle

    step(0)
    torch.cuda.synchronize()
    out0 = pool[0][3]
    oZ, ogdec, ogatt = oracle.dmv_log(md0, ma0, L0, trim=True)
    obest, oheads, _, _ = oracle.dmv_viterbi(md0, ma0, L0, trim=True)
    _, _, ogatt64 = oracle.dmv_log(md0, ma0, L0, trim=True, f64=True)
    gatt0 = out0.gattach.cpu().numpy()
    tol = np.maximum(1e-5, 2.0 ** -22 * np.abs(oZ.astype(np.float64))).reshape(-1, 1, 1, 1)  # 1e-5 or two ulps of |log Z|
    parity = {
        "heads_bit_exact": bool(np.array_equal(out0.heads.cpu().numpy(), oheads)),
        "best_bit_exact": bool(np.array_equal(out0.best.cpu().numpy(), obest)),
        "Z_max_rel": float(np.abs((out0.Z.cpu().numpy() - oZ) / oZ).max()),
        "marginal_max_abs_vs_f32_oracle": float(np.abs(gatt0 - ogatt).max()),
        "marginal_max_abs_vs_f64_oracle": float(np.abs(gatt0 - ogatt64).max()),
        "f32_oracle_max_abs_vs_f64_oracle": float(np.abs(ogatt - ogatt64).max()),
        "marginals_within_tol": bool((np.abs(gatt0 - ogatt) <= tol).all()),
    }
    if not (parity["heads_bit_exact"] and parity["best_bit_exact"] and parity["Z_max_rel"] < 1e-4
            and parity["marginals_within_tol"]):
        raise SystemExit(f"bench.py: parity gate failed on the timed configuration: {parity}")

    # ---- roofline denominators measured live (MUFU / FP32 issue rate are not in MEASURED_PEAKS.json) ----
    ms = ctypes.c_float()
    nops = ctypes.c_double()
    peaks = {}
    for name, fn in (("mufu", L_.vlgae_microbench_mufu), ("fp32", L_.vlgae_microbench_fp32)):
        best = 0.0
        for _ in range(3):
            check(fn(4000, ctypes.byref(ms), ctypes.byref(nops), None), "microbench")
            best = max(best, nops.value / (ms.value * 1e-3))
        peaks[name] = best

    # ---- timed region ----
    sampler = ClockSampler(local)
    sampler.start()
    for k in range(args.warmup):
        step(k)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.timed = True
    ev0.record()
    for k in range(args.steps):
        step(args.warmup + k)
    ev1.record()
    torch.cuda.synchronize()
    sampler.timed = False
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    elapsed_ms = ev0.elapsed_time(ev1)

    # ---- end to end: host buffers -> H2D -> kernels -> D2H, through the C ABI's host entry point ----
    pin = lambda a: torch.from_numpy(a).pin_memory()  # noqa: E731
    h_md, h_ma, h_L = pin(md0), pin(ma0), pin(L0)
    h_Z, h_best = torch.empty(B).pin_memory(), torch.empty(B).pin_memory()
    h_gatt = torch.empty((B, N, N, 2)).pin_memory()
    h_gdec = torch.empty((B, N, 2, 2, 2)).pin_memory()
    h_heads = torch.empty((B, N), dtype=torch.int64).pin_memory()
    stream = torch.cuda.current_stream().cuda_stream

    def e2e_step():
        check(L_.vlgae_dmv_parse_host(h_md.data_ptr(), h_ma.data_ptr(), h_L.data_ptr(), B, N, MASK_ZERO,
                                      h_Z.data_ptr(), h_gdec.data_ptr(), h_gatt.data_ptr(), h_best.data_ptr(),
                                      h_heads.data_ptr(), stream), "vlgae_dmv_parse_host")

    e2e_steps = max(10, min(args.steps, 200))
    for _ in range(3):
        e2e_step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()  # synchronises the stream itself (results are in host memory on return)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    sampler.stop()
    e2e_ok = bool(np.array_equal(h_heads.numpy(), oheads))
    h2d = md0.nbytes + ma0.nbytes + L0.nbytes
    d2h = h_Z.numel() * 4 + h_best.numel() * 4 + h_gatt.numel() * 4 + h_gdec.numel() * 4 + h_heads.numel() * 8

    # ---- alignment kernel, reported separately (rank 0 only) ----
    align = None
    if rank == 0 and not args.no_align:
        del pool, pmd, pma
        torch.cuda.empty_cache()
        align = alignment_leg(dev)

    # ---- max over ranks ----
    t = torch.tensor([elapsed_ms, e2e_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_s = float(t[0]), float(t[1])

    if rank == 0:
        wc = work_counts(L0)
        per_launch_s = elapsed_ms * 1e-3 / args.steps
        achieved = wc["mufu"] / per_launch_s
        value = B * world * args.steps / (elapsed_ms * 1e-3)
        cb = cpu_baseline(md0, ma0, L0) if world == 1 else None
        line = {
            "metric": "sentences/sec (inside+outside+Viterbi, len<=40)", "value": value, "unit": "sentences/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg2: 128 captions per GPU, len 4..40 ragged sorted desc (BASELINE.json "
                                   "configs[1]), DMV inside+outside+Viterbi in one launch",
                       "batch_per_gpu": B, "max_len": MAX_LEN, "parallelism": f"sentence-sharded x{world}, no collective",
                       "l2": f"inputs rotate through a pool of {pool_n} distinct batches "
                             f"({pool_n * step_bytes / 2**20:.0f} MiB > L2), no flush needed"},
            "e2e": {"value": B * world * e2e_steps / e2e_s, "unit": "sentences/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3,
                    "api": "vlgae_dmv_parse_host (pinned host buffers; zero-copy: the kernel pulls the potentials from and pushes the results to host memory over PCIe inside the timed call, then the stream is synchronised)",
                    "heads_bit_exact": e2e_ok},
            "gpu_launches": args.steps,
            "roofline": {"bound": "sfu", "achieved": achieved / 1e9, "peak": peaks["mufu"] / 1e9, "unit": "Gop/s",
                         "frac": achieved / peaks["mufu"],
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the ncu --set full capture
                         # summarised in profiles/r1_dmv_kernel.txt (outputs stay in L2 at capture time)
                         "traffic": 1.0045e6,
                         "kernel": "dmv_frontier_kernel<512,2> (frontier schedule: one thread per target cell, running logsumexp / arg-max state in registers, chart in shared memory)", "algorithmic_mufu_ops_per_launch": wc["mufu"],
                         "peak_source": "measured live: ex2.approx.f32 microbenchmark (vlgae_microbench_mufu)",
                         "fp32_frac": (wc["fp32"] / per_launch_s) / peaks["fp32"],
                         "fp32_peak_gops": peaks["fp32"] / 1e9,
                         "hbm_gbs": wc["hbm_bytes"] / per_launch_s / 1e9},
            "parity": parity,
            "clocks": sampler.summary(),
        }
        if cb is not None:
            line["cpu_baseline"] = cb
        if align is not None:
            line["alignment"] = align
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-align", action="store_true", help="skip the alignment-kernel leg")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps > 200:
            args.steps = 20  # the CPU arm's default: 20 bounded steps
            args.warmup = min(args.warmup, 2)
        return run_reference_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
