#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): per-kernel headline metrics + top stalled SASS."""
import csv, subprocess, sys, io
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = ["Kernel Name", "gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
for w in want:
    if w in idx:
        print(f"{w} [{units[idx[w]]}]: " + " | ".join(r[idx[w]] for r in data))
names = [r[idx["Kernel Name"]] for r in data]
for kn in dict.fromkeys(n.split("(")[0].split("<")[0].split()[-1] for n in names):
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kn}"], capture_output=True, text=True).stdout
    r2 = list(csv.reader(io.StringIO(src)))
    if len(r2) < 3:
        continue
    h2 = r2[1]; i2 = {h: i for i, h in enumerate(h2)}
    d2 = [r for r in r2[2:] if len(r) == len(h2) and r[i2["# Samples"]].strip().isdigit()]
    tot = sum(int(r[i2["# Samples"]]) for r in d2)
    print(f"\n== {kn}: {tot} samples, {len(d2)} SASS instructions; top {topn} by samples")
    for k in sorted(sorted(range(len(d2)), key=lambda k: -int(d2[k][i2["# Samples"]]))[:topn]):
        r = d2[k]
        print(f"  {k:5d} {r[i2['Source']].strip()[:58]:58s} samples={r[i2['# Samples']]:>7s} long={r[i2['stall_long_sb']]:>7s} short={r[i2['stall_short_sb']]:>6s} wait={r[i2['stall_wait']]:>6s} mio={r[i2['stall_mio']]:>6s} lg={r[i2['stall_lg']]:>6s}")
