"""Per-phase table of k_walk_chunks_fast from an ncu report captured with --import-source on.

    ncu -i report.ncu-rep --page source --csv > source.csv
    python tools/ncu_phase_table.py source.csv

The kernel's phases are separated by __syncwarp() (WARPSYNC.ALL in SASS; the compiler merges
some adjacent ones), so the SASS between two of them is one phase group. For each group: static
instructions, executed warp instructions and their share, average active lanes, share of the
warp-stall samples.
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
groups, cur = [], [0, 0, 0, 0]
for r in data:
    if len(r) <= ix["# Samples"]:
        continue
    if "WARPSYNC" in r[ix["Source"]]:
        groups.append(cur)
        cur = [0, 0, 0, 0]
    cur[0] += 1
    cur[1] += int(r[ix["Instructions Executed"]])
    cur[2] += int(r[ix["Thread Instructions Executed"]])
    cur[3] += int(r[ix["# Samples"]])
groups.append(cur)
tw = sum(g[1] for g in groups)
tt = sum(g[2] for g in groups)
ts = sum(g[3] for g in groups)
print(f"warp instructions {tw:.4e}  thread instructions {tt:.4e}  lanes/instruction {tt / tw:.2f}  samples {ts}")
print("group  static   warp-inst   share  lanes  stall-samples")
for k, g in enumerate(groups):
    print(f"{k:5d}  {g[0]:6d}  {g[1]:.3e}  {100 * g[1] / tw:5.1f}%  {g[2] / max(1, g[1]):5.1f}  {100 * g[3] / ts:5.1f}%")
