"""Regenerates tests/golden/* from the UNMODIFIED reference (oracle/_ref, built by
`make -C oracle ref` where /root/reference exists). Run from the repo root:

    python tests/golden/make_golden.py

Outputs
  c1.fa.gz        `test_fasta -s 1729 -l 100000 -d 0.01` (config 1 of BASELINE.json; the
                  reference's simulator, test/test_fasta.cxx)
  golden.json     reference results: 17-word models for c1 under every model, the printed CLI
                  matrix, estimates; ESA arrays (SA, LCP, CLD, FVC, cache digest) of the two
                  test_esa.c fixtures; get_match vectors; stress-set models.
"""
import ctypes as C
import gzip
import hashlib
import json
import random
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle  # noqa: E402
from conftest import ESA_FIXTURE_1, ESA_FIXTURE_2, stress_sequences  # noqa: E402

OUT = Path(__file__).resolve().parent


def main():
    assert oracle.ref_available(), "build oracle/_ref first"
    g = {}
    fa = oracle.run_test_fasta(1729, 100000, [0.01])
    g["c1_md5"] = hashlib.md5(fa).hexdigest()
    with gzip.GzipFile(OUT / "c1.fa.gz", "wb", mtime=0) as f:
        f.write(fa)
    seqs = [s for _, s in oracle.parse_fasta(fa)]
    g["c1"] = {}
    for model in oracle.MODELS:
        rows, _ = oracle.ref_rows(seqs, model)
        g["c1"][model] = rows.reshape(-1, 17).tolist()
    cli = {}
    for args in (["-t", "1"], ["-t", "1", "-m", "RAW"], ["-t", "1", "-m", "KIMURA"], ["-t", "1", "-m", "LOGDET"], ["-t", "1", "-v"], ["-t", "1", "-l"]):
        p = subprocess.run([str(oracle.REF_ANDI), *args], input=fa, capture_output=True, check=True)
        cli[" ".join(args)] = p.stdout.decode()
    g["c1_cli"] = cli
    R = oracle.ref()
    m01 = oracle.Model((C.c_uint32 * 16)(*g["c1"]["JC"][1][:16]), 100000)
    g["c1_estimates_m01"] = {k: getattr(R, "estimate_" + k)(C.byref(m01)) for k in ("RAW", "JC", "KIMURA", "LOGDET", "ANI")}
    h = oracle.RefEsaHandle(seqs[0])
    g["c1_threshold"] = h.threshold
    h.close()

    g["esa"] = []
    rng = random.Random(42)
    for s in (ESA_FIXTURE_1, ESA_FIXTURE_2):
        h = oracle.RefEsaHandle(s)
        ent = {"seq": s.decode(), "rs": h.rs.decode(), "threshold": h.threshold}
        for name in ("SA", "LCP", "FVC"):
            ent[name] = h.array(name).tolist()
        ent["CLD"] = h.array("CLD")[:-1].tolist()
        ent["cache_sha256"] = hashlib.sha256(h.array("cache").tobytes()).hexdigest()
        qs = [b"AAGACTGG", b"AATTAAAA", b"ACCGAGAA", b"AAAAAAAAAAAA", b"A", b"C", b"CT", b"!AAAAAAAAAAA"]
        for _ in range(200):
            ln = rng.choice([3, 8, 11, 15, 30])
            if rng.random() < 0.6:
                p = rng.randrange(0, len(s) - ln)
                q = bytearray(s[p : p + ln])
                if rng.random() < 0.5:
                    q[rng.randrange(ln)] = rng.choice(b"ACGT")
                qs.append(bytes(q))
            else:
                qs.append(bytes(rng.choice(b"ACGT") for _ in range(ln)))
        ent["queries"] = [q.decode() for q in qs]
        ent["get_match"] = [list(h.get_match(q, cached=False)[:3]) for q in qs]
        ent["get_match_cached"] = [list(h.get_match(q, cached=True)[:3]) for q in qs]
        g["esa"].append(ent)
        h.close()

    g["stress"] = {}
    for name, ss in stress_sequences().items():
        g["stress"][name] = {}
        for model in ("JC", "LOGDET"):
            rows, _ = oracle.ref_rows(ss, model)
            g["stress"][name][model] = rows.reshape(-1, 17).tolist()
        g["stress"][name]["sha256"] = hashlib.sha256(b"\n".join(ss)).hexdigest()

    (OUT / "golden.json").write_text(json.dumps(g, separators=(",", ":")))
    print("wrote", OUT / "golden.json", (OUT / "golden.json").stat().st_size, "bytes;", (OUT / "c1.fa.gz").stat().st_size)


if __name__ == "__main__":
    main()
