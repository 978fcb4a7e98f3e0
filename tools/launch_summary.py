"""Aggregate an ncu launch-list CSV (--metrics gpu__time_duration.sum) per kernel:
    python tools/launch_summary.py launches.csv > summary.txt     (kernel | launches | total_us | share | us per launch)"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
ix = {n: i for i, n in enumerate(rows[hdr])}
agg = collections.OrderedDict()
for r in rows[hdr + 2:]:
    try:
        k, v = r[ix["Kernel Name"]].split("(")[0][:72], float(r[ix["Metric Value"]].replace(",", ""))
    except (ValueError, IndexError):
        continue
    u = r[ix["Metric Unit"]]
    v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print("# kernel | launches | total_us | share | us per launch")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:72s} {a[0]:5d} {a[1]:10.1f} {100 * a[1] / tot:6.2f}% {a[1] / a[0]:9.1f}")
