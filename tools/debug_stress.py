"""Helper for `compute-sanitizer python tests/debug_sanitize.py` on a GPU box (not a test)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import numpy as np

import oracle
from andi_b200 import native
from conftest import stress_sequences

ctx = native.Context(0)
only = sys.argv[1:]
for name, seqs in stress_sequences().items():
    if only and name not in only:
        continue
    ctx.set_pool(seqs)
    got = ctx.dist_rows(model="JC")
    want = oracle.rows(seqs, "JC")
    print(name, "OK" if np.array_equal(got, want) else "MISMATCH", flush=True)
    if not np.array_equal(got, want):
        bad = np.argwhere((got != want).any(axis=2))
        print(bad[:5], got[tuple(bad[0])], want[tuple(bad[0])])
