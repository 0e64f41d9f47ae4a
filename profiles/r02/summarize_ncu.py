"""Compact table of one `ncu --set full` capture (`ncu -i X.ncu-rep --page raw --csv` output).
usage: python profiles/r02/summarize_ncu.py raw.csv [raw2.csv ...]   -> markdown on stdout"""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / CTA"),
    ("launch__occupancy_limit_registers", "CTA/SM limit (registers)"),
    ("launch__occupancy_limit_shared_mem", "CTA/SM limit (smem)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/TEX throughput"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput"),
    ("sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "FMA-heavy pipe"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "FMA-heavy pipe cycles"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe"),
    ("sm__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__warps_active.avg.per_cycle_active", "warps / scheduler"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / scheduler"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "of which bank conflicts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math pipe throttle / issue"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall MIO throttle / issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall LG throttle / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not selected / issue"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no instruction / issue"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar / issue"),
    ("smsp__average_warps_issue_stalled_drain_per_issue_active.ratio", "stall drain / issue"),
]

for f in sys.argv[1:]:
    rows = list(csv.reader(open(f)))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[h.index("Kernel Name")].split("(")[0]
        print("### `%s`  grid %s x block %s  (%s)\n" % (name, r[h.index("Grid Size")], r[h.index("Block Size")], f.split("/")[-1]))
        print("| metric | value |\n|---|---|")
        for k, label in KEYS:
            if k in h:
                i = h.index(k)
                v = r[i]
                try:
                    v = "%.4g" % float(v.replace(",", ""))
                except ValueError:
                    pass
                print("| %s | %s %s |" % (label, v, units[i]))
        print()
