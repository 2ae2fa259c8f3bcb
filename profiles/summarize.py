#!/usr/bin/env python
"""profiles/summarize.py -- turn raw ncu output (gpurun_out/) into the small text summaries committed here.

    python profiles/summarize.py launches gpurun_out/launches_r1.csv            > profiles/r1_launches.txt
    python profiles/summarize.py report   gpurun_out/prof_r1_dmv.ncu-rep [top]  > profiles/r1_dmv_kernel.txt
"""
import csv
import subprocess
import sys
from collections import defaultdict


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = defaultdict(list)
    order = []
    for r in rows[1:]:
        k = (r[ki], r[gi], r[bi])
        if k not in agg:
            order.append(k)
        agg[k].append(float(r[vi].replace(",", "")))
    total = sum(sum(v) for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none  ({path})")
    print(f"# {sum(len(v) for v in agg.values())} launches, {total / 1e3:.1f} us total device time (cold-cache, serialised)")
    print(f"{'share':>7} {'launches':>8} {'avg_ns':>10} {'min_ns':>10} {'max_ns':>10}  kernel  grid  block")
    for k in sorted(agg, key=lambda k: -sum(agg[k])):
        v = agg[k]
        print(f"{100 * sum(v) / total:6.1f}% {len(v):8d} {sum(v) / len(v):10.0f} {min(v):10.0f} {max(v):10.0f}  {k[0][:90]}  {k[1]}  {k[2]}")


RAW_KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "sm__cycles_elapsed.max", "sm__cycles_active.avg", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__warps_eligible.avg.per_cycle_active",
]


def report(path, top=25):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none --import-source on   ({path})")
    for v in rows[2:]:
        print(f"## kernel: {v[h.index('Kernel Name')][:100]}")
        for k in RAW_KEYS:
            for i, n in enumerate(h):
                if n == k:
                    print(f"{n:70s} {v[i]:>16s} {u[i]}")
        stall = [(float(v[i] or 0), n) for i, n in enumerate(h)
                 if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio")]
        print("warp stall reasons (warps per issue-active cycle):")
        for x, n in sorted(stall, reverse=True)[:8]:
            print(f"    {n[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {x:.3f}")
    src = subprocess.run([sys.executable, __file__.replace("summarize.py", "ncu_lines.py"), path, str(top)],
                         capture_output=True, text=True).stdout
    print("## hottest CUDA source lines (stall samples, warp instructions executed)")
    print(src)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        report(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
