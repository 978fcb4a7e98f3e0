"""Debug helper (GPU box): suffix array / LCP of the stress texts against the oracle, first mismatch printed."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import numpy as np

import oracle
from andi_b200 import native
from conftest import stress_sequences

ctx = native.Context(0)
groups = stress_sequences()
for name in ("subst", "repeat", "short", "lowent", "indel"):
    cases = [s for s in groups[name] if b"!" not in s]
    ctx.set_pool(cases)
    for k, s in enumerate(cases):
        o = oracle.OracleEsa(s)
        e = ctx.esa_build(k)
        got = e.download()
        for arr in ("SA", "LCP"):
            want = o.array(arr)
            bad = np.flatnonzero(got[arr] != want)
            if len(bad):
                j = int(bad[0])
                print(name, k, len(s), arr, "mismatches", len(bad), "first at", j, "got", got[arr][max(0, j - 2) : j + 4].tolist(),
                      "want", want[max(0, j - 2) : j + 4].tolist())
            else:
                print(name, k, len(s), arr, "ok")
        e.free(), o.close()
print(ctx.stats())
