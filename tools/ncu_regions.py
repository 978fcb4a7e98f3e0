"""Segment the SASS of one kernel of an ncu report (captured with --import-source on) into regions of
equal execution count and print, per region, its share of the executed warp instructions, the
average number of active lanes and its share of the stall samples.

    ncu -i report.ncu-rep --page source --csv --print-source sass > source.csv
    python tools/ncu_regions.py source.csv [kernel-name-substring] [listing-out.txt]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
kern, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        kern.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and len(r) > 7:
        cur["rows"].append(r)
k = [x for x in kern if want in x["name"]][0]
ix = {h: i for i, h in enumerate(k["hdr"])}
base = int(k["rows"][0][0], 16)
out = []
for r in k["rows"]:
    out.append((int(r[0], 16) - base, int(r[ix["Instructions Executed"]]), int(r[ix["Thread Instructions Executed"]]),
                int(r[ix["# Samples"]]), r[1].strip()))
tw, tt, ts = sum(o[1] for o in out), sum(o[2] for o in out), sum(o[3] for o in out)
print(f"{k['name'][:40]}: warp instructions {tw:.4e}, thread instructions {tt:.4e}, lanes/instruction {tt / tw:.2f}")
if len(sys.argv) > 3:
    open(sys.argv[3], "w").write("\n".join(f"{a:05x} {w / 1e6:8.2f}M {t / max(1, w):5.1f} {100 * s / ts:5.2f}% {src}" for a, w, t, s, src in out))
seg = []
for a, w, t, s, src in out:
    if seg and abs(seg[-1]["w"] - w) <= 0.02 * max(w, seg[-1]["w"]):
        g = seg[-1]
        g["n"] += 1; g["tw"] += w; g["tt"] += t; g["s"] += s; g["end"] = a
    else:
        seg.append({"start": a, "end": a, "w": w, "n": 1, "tw": w, "tt": t, "s": s})
print("start  end    static  executions(M)  share  lanes  stall-samples")
for g in seg:
    if g["tw"] / tw > 0.004:
        print(f"{g['start']:05x} {g['end']:05x} {g['n']:6d} {g['w'] / 1e6:13.2f} {100 * g['tw'] / tw:6.1f}% {g['tt'] / g['tw']:5.1f} {100 * g['s'] / ts:8.1f}%")
