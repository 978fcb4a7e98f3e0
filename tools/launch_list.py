"""Workload for the ncu launch-list pass (profiles/*_launch_list_*): a small pool made on the host
with numpy, so that every kernel in the list is one of this library's.

    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
        --log-file gpurun_out/launches.csv python tools/launch_list.py [genomes] [length] [rows]

Per-launch times under ncu are cold-cache and serialised: use them for each kernel's SHARE of a
subject, not as absolute numbers (bench.py times the real thing with CUDA events).
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np

from andi_b200 import native, synth

genomes = int(sys.argv[1]) if len(sys.argv) > 1 else 64
length = int(sys.argv[2]) if len(sys.argv) > 2 else 2_100_000
rows = int(sys.argv[3]) if len(sys.argv) > 3 else 4

rng = np.random.default_rng(1)
seqs = synth.star_phylogeny(genomes, length, list(rng.uniform(0.005, 0.02, size=genomes)), seed=3085)
ctx = native.Context(0)
ctx.set_pool(seqs)
M = ctx.dist_rows(s_begin=0, s_end=rows, model="JC")
print("rows", M.shape, "cell[0,1]", M[0, 1][:4].tolist(), ctx.stats())
