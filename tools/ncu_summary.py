#!/usr/bin/env python3
"""Summarise ncu captures brought back in gpurun_out/ into small text files for profiles/.

    python tools/ncu_summary.py rep  gpurun_out/prof_x.ncu-rep   > profiles/x_summary.txt
    python tools/ncu_summary.py list gpurun_out/launches.csv      > profiles/launches.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__inst_executed_op_branch.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print("# ncu --set full summary of %s" % path)
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("kernel: %s" % name)
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("  %-82s %16s %s" % (k, r[i], units[i]))


def launches(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= mv:
            continue
        name = r[kn].split("(")[0][:70]
        v = float(r[mv].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[mu], 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print("# ncu launch list (gpu__time_duration.sum, --clock-control none) of %s" % path)
    print("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes")
    print("%-72s %6s %12s %7s" % ("kernel", "n", "total ms", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-72s %6d %12.3f %6.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
    print("%-72s %6s %12.3f" % ("TOTAL", "", tot))


if __name__ == "__main__":
    {"rep": rep, "list": launches}[sys.argv[1]](sys.argv[2])
