"""GPU box: the andi command line end to end on FASTA files of the C4 shape: writes
`genomes` files of 2.1 Mbp, runs andi_b200/andi on them (files read by all cores) and prints the
wall time next to the same run with OMP_NUM_THREADS=1 (the reference reads its files serially)."""
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np

from andi_b200 import synth

genomes = int(sys.argv[1]) if len(sys.argv) > 1 else 256
root = Path(__file__).resolve().parent.parent
d = Path(tempfile.mkdtemp(prefix="andi_cli_"))
rng = np.random.default_rng(7)
base = synth.base_genome(2_100_000, 3085)
names = []
for k in range(genomes):
    s = synth.ACGT[synth.mutate(base, float(rng.uniform(0.005, 0.02)), 100 + k)]
    lines = np.full((30000, 71), ord("\n"), np.uint8)
    lines[:, :70] = s.reshape(30000, 70)
    p = d / f"g{k:04d}.fa"
    with open(p, "wb") as f:
        f.write(b">g%04d synthetic\n" % k)
        f.write(lines.tobytes())
    names.append(str(p))
fof = d / "files.txt"
fof.write_text("\n".join(names) + "\n")
total = sum(os.path.getsize(n) for n in names)
for threads in ("16", "1"):
    env = dict(os.environ, OMP_NUM_THREADS=threads)
    t0 = time.perf_counter()
    r = subprocess.run([str(root / "andi_b200" / "andi"), "--file-of-filenames", str(fof)], capture_output=True, env=env)
    dt = time.perf_counter() - t0
    print(f"andi on {genomes} files ({total / 1e6:.0f} MB FASTA), {threads} host thread(s) for the ingest: {dt:.2f} s wall, rc {r.returncode}, "
          f"{len(r.stdout)} bytes of PHYLIP", flush=True)
