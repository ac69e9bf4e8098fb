"""Turns an .ncu-rep (ncu --set full of ONE kernel launch) into the text summary kept under profiles/: selected raw
metrics, the executed-instruction mix per warp from the source page, stall samples by reason.
usage: python profiles/ncu_summary.py <file.ncu-rep> <warps-per-unit-divisor> [title]"""
import collections
import csv
import subprocess
import sys

rep, div = sys.argv[1], float(sys.argv[2])
title = sys.argv[3] if len(sys.argv) > 3 else rep
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h, u, v = raw[0], raw[1], raw[2]
print(f"# {title}\n# source: ncu --set full --clock-control none --import-source on, one launch; file {rep.split('/')[-1]}")
print("Kernel:", v[h.index("Kernel Name")] if "Kernel Name" in h else "?")
for k in KEYS:
    if k in h:
        print(f"{k:95s} {u[h.index(k)]:>16s} {v[h.index(k)]}")
src = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout.splitlines()))
hi = [i for i, r in enumerate(src) if r and r[0] == "Address"]
hdr = src[hi[0]]
data = [r for r in src[hi[0] + 1:(hi[1] if len(hi) > 1 else len(src))] if len(r) == len(hdr)]
ix = {k: i for i, k in enumerate(hdr)}
tot, byop, stall = 0, collections.Counter(), collections.Counter()
for r in data:
    ex = int(r[ix["Instructions Executed"]])
    tot += ex
    s = r[ix["Source"]].split()
    op = (s[1] if s[0].startswith("@") else s[0]).split(".")[0]
    byop[op] += ex
    for k in hdr:
        if k.startswith("stall_") and "(" not in k:
            stall[k] += int(r[ix[k]])
fp = sum(byop[k] for k in ("DFMA", "DMUL", "DADD"))
print(f"\nexecuted warp-instructions per unit (divisor {div:g}): total {tot / div:.1f}, FP64 {fp / div:.1f}, other {(tot - fp) / div:.1f}")
print("  " + ", ".join(f"{op} {c / div:.2f}" for op, c in byop.most_common(24)))
ts = sum(stall.values())
print("warp-state samples: " + ", ".join(f"{k[6:]} {100 * c / ts:.1f}%" for k, c in stall.most_common(9)))
