#!/usr/bin/env python3
"""Key metrics per kernel launch from an `ncu --page raw --csv` export: python tools/ncu_brief.py gpurun_out/x_full_raw.csv [name filter]"""
import csv
import sys

WANT = [("gpu__time_duration.sum", "us"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem"), ("launch__occupancy_limit_shared_mem", "occ smem"),
        ("launch__occupancy_limit_registers", "occ regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__issue_active.avg.pct", "issue %"), ("smsp__inst_executed.sum", "warp inst"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/inst"),
        ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"), ("lts__t_bytes.sum", "l2 bytes"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem conflicts"),
        ("l1tex__data_pipe_lsu_wavefronts.sum", "lsu wavefronts"), ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu pipe %"),
        ("l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active", "lsu wb %"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short sb"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long sb"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no inst"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe %"), ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu pipe %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm thr %"), ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "mem thr %")]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
flt = sys.argv[2] if len(sys.argv) > 2 else ""
seen = set()
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    if flt not in name or name in seen:
        continue
    seen.add(name)
    print("==", name)
    for k, lab in WANT:
        if k in idx:
            print("  %-16s %s %s" % (lab, r[idx[k]], units[idx[k]]))
