#!/usr/bin/env python
"""Key metrics of the kernels in an .ncu-rep (raw page): tools/ncu_keys.py file.ncu-rep [kernel regex]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]; pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[0]
keys = ["gpu__time_duration.sum", "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tc.sum", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.avg.per_cycle_active", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct", "smsp__warp_issue_stalled_sleeping_per_warp_active.pct",
        "smsp__warp_issue_stalled_membar_per_warp_active.pct", "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct", "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct"]
ki = h.index("Kernel Name")
for r in rows[2:]:
    if pat and not pat.search(r[ki]): continue
    print("==", r[ki][:70])
    for k in h:
        if any(k == q or k.startswith(q) for q in keys):
            print(f"   {k:90s} {r[h.index(k)]}")
