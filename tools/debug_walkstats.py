"""Instrumentation helper (not a test): lane-trip counts per phase of k_walk_chunks_fast.
Run on a GPU box with ANDI_B200_LIB=andi_b200/_whatif/libandi_stats.so (built with
-DANDI_WALK_STATS)."""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

import bench
from andi_b200 import native

g, ln = int(sys.argv[1]) if len(sys.argv) > 1 else 400, 2_100_000
lo, hi = (float(sys.argv[2]), float(sys.argv[3])) if len(sys.argv) > 3 else (0.005, 0.02)
chars, offsets, lens, d = bench.make_pool_device(g, ln, lo, hi, 3085, torch.device("cuda", 0))
ctx = native.Context(0)
ctx.set_pool_device(chars.data_ptr(), offsets, lens)
out = torch.empty((1, g, 17), dtype=torch.int32, device="cuda")
L = native.load()
buf = (C.c_ulonglong * 16)()
ctx.dist_rows_device(out.data_ptr(), 0, 1)
L.andi_debug_walk_stats(buf, 1)
ctx.dist_rows_device(out.data_ptr(), 1, 2)
L.andi_debug_walk_stats(buf, 1)
v = list(buf)
names = ["BEGIN", "CMP1", "DIR", "CAND", "CMP2", "SLOW", "DECIDE", "COLS", "CMP1cand", "CMP2cand", "DECIDEfound", "-", "-", "-", "lanes_active", "warp_trips"]
trips = v[15]
print("pairs", g - 1, "warp trips", trips, "active lane-trips", v[14], "= %.1f lanes/trip" % (v[14] / trips))
for n, x in zip(names, v):
    if n != "-":
        print("%-12s lane-trips %12d  per warp-trip %5.2f  per DECIDE %5.2f" % (n, x, x / trips, x / max(1, v[6])))
print("steps(DECIDE) per pair: %.0f ; trips per pair-lane: %.0f" % (v[6] / (g - 1), v[14] / (g - 1)))
