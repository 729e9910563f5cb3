#!/usr/bin/env python
"""Keep the columns the profiles/ summaries quote from an `ncu -i report --page raw --csv` dump (stdin -> stdout)."""
import csv
import sys

KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "sm__cycles_active.avg",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]

rows = list(csv.reader(l for l in sys.stdin if not l.startswith("==")))
hdr = rows[0]
idx = [hdr.index(k) for k in KEEP if k in hdr]
w = csv.writer(sys.stdout)
for r in rows:
    if len(r) >= len(hdr):
        w.writerow([r[i][:60] if hdr[i] == "Kernel Name" else r[i] for i in idx])
