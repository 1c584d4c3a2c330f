#!/usr/bin/env python
"""Compact text summary of an `ncu --page raw --csv` export (and optionally the source page):
   python tools/ncu_summary.py raw.csv [src.csv] > profiles/<name>.txt"""
import collections
import csv
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn = hdr.index("Kernel Name") if "Kernel Name" in hdr else None
    for r in data:
        print("kernel:", r[kn][:110] if kn is not None else "?")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("  %-82s %-10s %s" % (k, units[i], r[i]))
    if len(sys.argv) > 2:
        rows = list(csv.reader(open(sys.argv[2])))
        hdr, data = rows[1], rows[2:]
        ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
        tot = sum(int(r[ie]) for r in data)
        ts = sum(int(r[isamp]) for r in data)
        ops = collections.Counter()
        for r in data:
            t = r[ia].strip().split()
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            ops[op] += int(r[ie])
        print("warp instructions executed: %d ; stall samples: %d" % (tot, ts))
        print("top opcodes by executed count:")
        for op, c in ops.most_common(16):
            print("  %-10s %6.2f %%" % (op, 100.0 * c / tot))
        sass = " ".join(r[ia] for r in data)
        for pat in ("UBLKCP", "SYNCS.PHASECHK", "LDG.E.ENL2.256", "MUFU.RSQ64H", "MUFU.RCP64H", "DFMA", "HMMA", "UTC"):
            print("  SASS contains %-16s %s" % (pat, pat in sass))


if __name__ == "__main__":
    main()
