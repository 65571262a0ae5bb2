#!/usr/bin/env python
"""Summarise .ncu-rep files (ncu --page raw --csv) into one table: the numbers the roofline needs."""
import csv
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "dur_us"), ("dram__bytes_read.sum", "rd_MB"), ("dram__bytes_write.sum", "wr_MB"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__inst_executed_pipe_fp64.sum", "fp64_inst"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%"),
        ("smsp__inst_executed.sum", "inst"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wf"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank_conf"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"), ("launch__registers_per_thread", "regs"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "st_sleep"),
        ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "st_membar"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_notsel"),
        ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "st_disp"),
        ("smsp__average_warps_issue_stalled_drain_per_issue_active.ratio", "st_drain"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "st_branch"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "st_noinst"),
        ("smsp__average_warps_issue_stalled_selected_per_issue_active.ratio", "st_sel")]
for f in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"== {f}: {name[:70]}")
        line = []
        for k, short in KEYS:
            if k in hdr:
                v = r[hdr.index(k)]
                u = units[hdr.index(k)]
                try:
                    x = float(v.replace(",", ""))
                    if u == "byte": x /= 1e6
                    if u == "Mbyte": pass
                    if u == "Gbyte": x *= 1e3
                    if u == "ns": x /= 1e3
                    if u == "ms": x *= 1e3
                    v = f"{x:.4g}"
                except ValueError:
                    pass
                line.append(f"{short}={v}")
        print("  " + "  ".join(line))
