#!/usr/bin/env python
"""Summarise ncu outputs into small text files for profiles/ (run here, no GPU needed).

  python tools/ncu_summary.py launches gpurun_out/X_launches.csv  > profiles/rNN_launches.txt
  python tools/ncu_summary.py full     gpurun_out/X_prof.ncu-rep  > profiles/rNN_<kernel>_full.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio"]


def launches(path):
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[h]
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) < len(H):
            continue
        d = dict(zip(H, r))
        v = float(d["Metric Value"].replace(",", ""))
        a = agg.setdefault(d["Kernel Name"][:90], [0, 0.0, d["Grid Size"], d["Block Size"]])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    print("# source: %s ; total %.3f ms over %d launches" % (path, tot / 1e6, sum(a[0] for a in agg.values())))
    print("%-92s %6s %12s %8s  %s" % ("kernel", "n", "total_us", "share", "grid/block (last)"))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-92s %6d %12.1f %8.4f  %s / %s" % (k, a[0], a[1] / 1e3, a[1] / tot, a[2], a[3]))


def full(path):
    out = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(out.splitlines()))
    H, U = rows[0], rows[1]
    print("# ncu --set full --clock-control none ; source: %s" % path)
    for V in rows[2:]:
        d = dict(zip(H, zip(U, V)))
        print("kernel: %s" % d["Kernel Name"][1])
        for k in KEYS:
            if k in d:
                print("  %-90s %16s %s" % (k, d[k][1], d[k][0]))
        rd = d.get("dram__bytes_read.sum")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
