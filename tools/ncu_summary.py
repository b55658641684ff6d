#!/usr/bin/env python
"""Summarise .ncu-rep files (read with `ncu -i ... --page raw --csv`), or that page saved as .csv, as a markdown table.

usage: tools/ncu_summary.py [--kernel REGEX] [--row N] label=path.ncu-rep [label=path ...]

One column per report (the N-th profiled launch matching REGEX, default the first).
The metric list is the one DESIGN.md argues from; extend METRICS to add more.
"""
import csv
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size",
    "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__warps_eligible.avg.per_cycle_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "local_load_requests", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
]


def load(path, pat, nth):
    if path.endswith(".csv"):          # `ncu -i x.ncu-rep --page raw --csv` made on the GPU box
        txt = open(path).read()
    else:
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                             text=True).stdout
    rows = [r for r in csv.reader(txt.splitlines()) if r and not r[0].startswith("==")]
    hdr, units = rows[0], rows[1]
    kcol = hdr.index("Kernel Name")
    sel = [r for r in rows[2:] if re.search(pat, r[kcol])]
    r = sel[nth]
    return r[kcol], {h: (v, u) for h, v, u in zip(hdr, r, units)}


def main():
    args = sys.argv[1:]
    pat, nth = ".", 0
    while args and args[0].startswith("--"):
        if args[0] == "--kernel":
            pat = args[1]
        elif args[0] == "--row":
            nth = int(args[1])
        args = args[2:]
    cols = []
    for a in args:
        label, path = a.split("=", 1)
        name, d = load(path, pat, nth)
        cols.append((label, name, d))
    print("| metric | " + " | ".join(c[0] for c in cols) + " |")
    print("|---|" + "---|" * len(cols))
    print("| kernel | " + " | ".join("`" + c[1].split("(")[0][:70] + "`" for c in cols) + " |")
    for m in METRICS:
        vals = [c[2].get(m) for c in cols]
        if all(v is None for v in vals):
            continue
        unit = next(v[1] for v in vals if v is not None)
        print(f"| `{m}` [{unit}] | " + " | ".join("" if v is None else v[0] for v in vals) + " |")


if __name__ == "__main__":
    main()
