"""Synthetic genomes in the shape of the reference's simulator (test/test_fasta.cxx:16-118):
a uniform iid ACGT base genome and, per sequence, exactly round(len * p) substitutions at
distinct uniform positions, each to one of the three other bases. The random streams are
numpy's, not libstdc++'s, so the sequences differ from test_fasta's for the same seed; the
committed golden fixtures under tests/golden/ hold genuine test_fasta output.
"""
from __future__ import annotations

import math

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def jc_to_raw(d: float) -> float:
    """test_fasta.cxx:52-58: evolutionary distance -> expected raw divergence."""
    return 0.75 - 0.75 * math.exp(-(4.0 / 3.0) * d)


def base_genome(length: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.integers(0, 4, size=length, dtype=np.uint8)


def mutate(codes: np.ndarray, p: float, seed: int) -> np.ndarray:
    """Exactly round(len*p) substitutions (codes are 0..3)."""
    rng = np.random.default_rng(seed)
    n = codes.shape[0]
    k = int(round(n * p))
    out = codes.copy()
    if k:
        pos = rng.choice(n, size=k, replace=False)
        out[pos] = (out[pos] + rng.integers(1, 4, size=k, dtype=np.uint8)) & 3
    return out


def star_phylogeny(n_seqs: int, length: int, divergences, seed: int, raw: bool = True):
    """n_seqs genomes descending from one base genome; divergences[k] is the raw (or, with
    raw=False, Jukes-Cantor) distance of genome k from the base. Returns list of bytes."""
    base = base_genome(length, seed)
    out = []
    for k in range(n_seqs):
        p = divergences[k] if raw else jc_to_raw(divergences[k])
        out.append(ACGT[mutate(base, p, seed + 1 + k)].tobytes())
    return out


def with_indels(seq: bytes, n_events: int, max_len: int, seed: int) -> bytes:
    """Parity-stress helper: random insertions/deletions so diagonals change."""
    rng = np.random.default_rng(seed)
    s = bytearray(seq)
    for _ in range(n_events):
        pos = int(rng.integers(0, max(1, len(s))))
        ln = int(rng.integers(1, max_len + 1))
        if rng.random() < 0.5:
            del s[pos : pos + ln]
        else:
            s[pos:pos] = ACGT[rng.integers(0, 4, size=ln)].tobytes()
    return bytes(s)


def join_contigs(seq: bytes, n_contigs: int, seed: int) -> bytes:
    """Cut a genome into contigs and glue them with '!' like dsa_join (src/sequence.c:78-125)."""
    rng = np.random.default_rng(seed)
    cuts = sorted(set(int(x) for x in rng.integers(1, len(seq) - 1, size=max(0, n_contigs - 1))))
    parts, last = [], 0
    for c in cuts:
        parts.append(seq[last:c])
        last = c
    parts.append(seq[last:])
    return b"!".join(parts)


def config_divergences(n_seqs: int, lo: float, hi: float, seed: int):
    rng = np.random.default_rng(seed ^ 0x5EED)
    return list(rng.uniform(lo, hi, size=n_seqs))
