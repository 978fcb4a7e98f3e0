"""Text summary of an ncu report (captured with --set full), one block per kernel launch: the
numbers DESIGN.md and bench.py quote.   python tools/ncu_summary.py report.ncu-rep > profiles/xyz.txt"""
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
WANT = [
    ("gpu__time_duration.sum", "Duration"),
    ("launch__grid_size", "Grid Size"), ("launch__block_size", "Block Size"),
    ("launch__registers_per_thread", "Registers Per Thread"),
    ("launch__shared_mem_per_block_static", "Static Shared Memory Per Block"),
    ("launch__shared_mem_per_block_dynamic", "Dynamic Shared Memory Per Block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "Achieved Occupancy"),
    ("sm__maximum_warps_per_active_cycle_pct", "Theoretical Occupancy"),
    ("smsp__inst_executed.sum", "Executed Warp Instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "Avg. Active Threads Per Warp"),
    ("smsp__thread_inst_executed_per_inst_executed.pct", None),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "Issue Slots Busy"),
    ("smsp__average_warp_latency_per_inst_issued.ratio", "Warp Cycles Per Issued Instruction"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "  stalled: long scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "  stalled: wait"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "  stalled: short scoreboard"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "  stalled: branch resolving"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "  stalled: not selected"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "  stalled: LG throttle"),
    ("smsp__sass_average_branch_targets_threads_uniform.pct", "Branch Efficiency"),
    ("dram__bytes_read.sum", "DRAM bytes read"), ("dram__bytes_write.sum", "DRAM bytes written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM Throughput"),
    ("dram__bytes_read.sum.per_second", "DRAM read rate"), ("dram__bytes_write.sum.per_second", "DRAM write rate"),
    ("l1tex__t_sector_hit_rate.pct", "L1/TEX Hit Rate"), ("lts__t_sector_hit_rate.pct", "L2 Hit Rate"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/TEX Cache Throughput"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 Cache Throughput"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "Compute (SM) Throughput"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe utilisation"),
]
for r in rows[2:]:
    print("==", r[h.index("Kernel Name")][:120])
    for key, label in WANT:
        if key in h and label:
            i = h.index(key)
            print(f"   {label:42s} {units[i]:>16s} {r[i]:>20s}")
    print()
