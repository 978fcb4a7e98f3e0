"""CPU tests of the C host code behind the command line (andi_b200/host): estimators against
the oracle (itself pinned to the reference), FASTA ingest, join mode."""
import ctypes as C
import gzip
from pathlib import Path

import numpy as np
import pytest

import oracle
from andi_b200 import native

ROOT = Path(__file__).resolve().parent.parent
G = ROOT / "tests" / "golden"


class HostSeq(C.Structure):
    _fields_ = [("S", C.c_char_p), ("len", C.c_size_t), ("name", C.c_char_p)]


class HostSeqs(C.Structure):
    _fields_ = [("data", C.POINTER(HostSeq)), ("size", C.c_size_t), ("capacity", C.c_size_t)]


@pytest.fixture(scope="module")
def host():
    L = C.CDLL(str(ROOT / "andi_b200" / "libandi_host.so"))
    L.model_estimate.restype = C.c_double
    L.model_estimate.argtypes = [C.POINTER(native.Model), C.c_int]
    L.model_coverage.restype = C.c_double
    L.model_coverage.argtypes = [C.POINTER(native.Model)]
    L.model_average.restype = native.Model
    L.model_average.argtypes = [C.POINTER(native.Model), C.POINTER(native.Model)]
    L.fasta_read.argtypes = [C.c_char_p, C.POINTER(HostSeqs), C.POINTER(C.c_int)]
    L.fasta_read_join.argtypes = [C.c_char_p, C.POINTER(HostSeqs), C.POINTER(C.c_int)]
    L.seqs_init.argtypes = [C.POINTER(HostSeqs)]
    L.seqs_free.argtypes = [C.POINTER(HostSeqs)]
    L.host_rng_new.restype = C.c_void_p
    L.host_rng_new.argtypes = [C.c_ulong]
    L.host_model_bootstrap.restype = native.Model
    L.host_model_bootstrap.argtypes = [C.c_void_p, native.Model]
    L.host_rng_binomial.restype = C.c_uint32
    L.host_rng_binomial.argtypes = [C.c_void_p, C.c_double, C.c_uint32]
    return L


def test_estimators_match_oracle(host):
    rng = np.random.default_rng(3)
    for trial in range(300):
        scale = int(rng.choice([10, 1000, 100000, 2000000]))
        diag = rng.integers(scale // 2, scale, size=4)
        off = rng.integers(0, max(2, scale // rng.choice([10, 50, 1000])), size=12)
        counts = np.zeros(16, np.uint32)
        counts[[0, 5, 10, 15]] = diag
        counts[[1, 2, 3, 4, 6, 7, 8, 9, 11, 12, 13, 14]] = off
        m = native.Model((C.c_uint32 * 16)(*counts), int(counts.sum()) + 7)
        om = oracle.Model((C.c_uint32 * 16)(*counts), int(counts.sum()) + 7)
        for name, mid in native.MODELS.items():
            got = host.model_estimate(C.byref(m), mid)
            want = oracle.lib().orc_estimate(C.byref(om), mid)
            assert (np.isnan(got) and np.isnan(want)) or got == pytest.approx(want, rel=1e-12, abs=0.0), (name, counts)
        assert host.model_coverage(C.byref(m)) == oracle.lib().orc_model_coverage(C.byref(om))
    tiny = native.Model((C.c_uint32 * 16)(1, 0, 1), 10)
    assert np.isnan(host.model_estimate(C.byref(tiny), 0))  # src/model.c:87-89: nucl <= 3


def test_fasta_reader_matches_fixture(host, tmp_path):
    fa = gzip.open(G / "c1.fa.gz").read()
    p = tmp_path / "c1.fa"
    p.write_bytes(fa)
    v = HostSeqs()
    host.seqs_init(C.byref(v))
    flags = C.c_int(0)
    assert host.fasta_read(str(p).encode(), C.byref(v), C.byref(flags)) == 0
    want = oracle.parse_fasta(fa)
    assert v.size == len(want) == 2
    for k, (name, seq) in enumerate(want):
        assert v.data[k].name.decode() == name and v.data[k].len == len(seq)
        assert C.string_at(v.data[k].S, v.data[k].len) == seq
    assert flags.value == 0
    host.seqs_free(C.byref(v))


def test_fasta_normalises_and_joins(host, tmp_path):
    p = tmp_path / "dir.with.dots" / "genome.v2.fasta"
    p.parent.mkdir()
    p.write_text(">c1 first contig\nacgtNNAC\nGT\n\n>c2\nTTTT-*\n>c3 x\nGGxG\n")
    v = HostSeqs()
    host.seqs_init(C.byref(v))
    flags = C.c_int(0)
    assert host.fasta_read(str(p).encode(), C.byref(v), C.byref(flags)) == 0
    assert [C.string_at(v.data[k].S) for k in range(3)] == [b"ACGTACGT", b"TTTT", b"GGG"]
    assert flags.value & 8  # HF_NON_ACGT
    host.seqs_free(C.byref(v))
    host.seqs_init(C.byref(v))
    assert host.fasta_read_join(str(p).encode(), C.byref(v), C.byref(flags)) == 0
    assert v.size == 1 and C.string_at(v.data[0].S) == b"ACGTACGT!TTTT!GGG"  # src/sequence.c:78-125
    assert v.data[0].name == b"genome"  # src/io.c:176-186: basename up to the first dot
    host.seqs_free(C.byref(v))


def test_fasta_errors_are_soft(host, tmp_path):
    for text in ("", "ACGT\n", ">\nACGT\n", ">name only"):
        p = tmp_path / "bad.fa"
        p.write_text(text)
        v = HostSeqs()
        host.seqs_init(C.byref(v))
        flags = C.c_int(0)
        assert host.fasta_read(str(p).encode(), C.byref(v), C.byref(flags)) == 1
        assert flags.value & 256  # HF_SOFT_ERROR
        host.seqs_free(C.byref(v))


def test_bootstrap_is_a_multinomial_resample(host):
    counts = [24436, 80, 94, 70, 83, 24432, 77, 85, 78, 87, 24406, 88, 84, 85, 81, 25720]
    m = native.Model((C.c_uint32 * 16)(*counts), 100000)
    rng = host.host_rng_new(42)
    draws = np.array([list(host.host_model_bootstrap(rng, m).counts) for _ in range(400)], dtype=np.float64)
    assert (draws.sum(axis=1) == sum(counts)).all()  # src/model.c:222-232: N is preserved
    mean, want = draws.mean(axis=0), np.array(counts, dtype=np.float64)
    assert np.all(np.abs(mean - want) < 5 * np.sqrt(want) / np.sqrt(400) + 1.0)


@pytest.mark.parametrize("n,p", [(7, 0.3), (40, 0.02), (25, 0.5), (200, 0.1), (1000, 0.031), (1000, 0.47), (5000, 0.9),
                                 (100000, 0.0004), (100000, 0.25), (2100000, 0.0125), (2100000, 0.245), (4000000000, 1e-6)])
def test_binomial_sampler_is_exact(host, n, p):
    """The bootstrap's binomial draws (inversion below a mean of 30, BTPE above) against the exact
    binomial pmf: chi-square goodness of fit over the bins that carry the mass, plus mean and
    variance. (The reference draws the same distribution through gsl_ran_binomial; its random
    stream cannot be reproduced without GSL, so the distribution is what is pinned here.)"""
    from scipy import stats

    draws = 40000
    rng = host.host_rng_new(n % 1000 + int(p * 1e6))
    x = np.array([host.host_rng_binomial(rng, p, n) for _ in range(draws)], dtype=np.int64)
    assert x.min() >= 0 and x.max() <= n
    mean, var = n * p, n * p * (1 - p)
    assert abs(x.mean() - mean) < 5 * np.sqrt(var / draws)
    assert abs(x.var() - var) < 6 * var * np.sqrt(2.0 / draws) + 5 * np.sqrt(var / draws)
    # bins: equal-probability classes from the exact quantiles (at most 60 of them)
    dist = stats.binom(n, p)
    edges = np.unique(dist.ppf(np.linspace(0, 1, 61)[1:-1]).astype(np.int64))
    cdf = np.concatenate([[0.0], dist.cdf(edges), [1.0]])
    expect = np.diff(cdf) * draws
    observed = np.bincount(np.searchsorted(edges, x, side="left"), minlength=len(edges) + 1)
    keep = expect > 5
    chi2 = float((((observed - expect) ** 2) / np.where(keep, expect, 1.0))[keep].sum())
    dof = int(keep.sum()) - 1
    assert dof >= 1
    assert chi2 < stats.chi2.ppf(1 - 1e-6, dof), (n, p, chi2, dof)


def test_bootstrap_cell_marginals(host):
    """model_bootstrap (src/model.c:222-232): every cell of the resampled matrix is Binomial(N, c/N),
    at small and at genome-sized counts; N is preserved."""
    from scipy import stats

    for counts in ([3, 0, 1, 0, 2, 5, 0, 0, 1, 0, 4, 0, 0, 0, 0, 6],
                   [524000, 2100, 2300, 1900, 2050, 518000, 2250, 1980, 2010, 2150, 530000, 1890, 2240, 2020, 2060, 526000]):
        m = native.Model((C.c_uint32 * 16)(*counts), sum(counts))
        rng = host.host_rng_new(7)
        reps = 3000
        draws = np.array([list(host.host_model_bootstrap(rng, m).counts) for _ in range(reps)], dtype=np.float64)
        assert (draws.sum(axis=1) == sum(counts)).all()
        N = sum(counts)
        for k, c in enumerate(counts):
            pk = c / N
            if c == 0:
                assert (draws[:, k] == 0).all()
                continue
            var = N * pk * (1 - pk)
            assert abs(draws[:, k].mean() - c) < 5 * np.sqrt(var / reps) + 1e-9, (k, c)
            assert abs(draws[:, k].var() - var) < 6 * var * np.sqrt(2.0 / reps) + 5 * np.sqrt(var / reps), (k, c)
        # two cells are negatively correlated: cov = -N p_i p_j
        cov = np.cov(draws[:, 0], draws[:, 5])[0, 1]
        want = -N * (counts[0] / N) * (counts[5] / N)
        assert abs(cov - want) < 6 * np.sqrt(counts[0] * counts[5]) / np.sqrt(reps) + 1.0


def test_host_packer_layout_and_counts():
    """host_pack.c (the opt-in host-side 2-bit packing of andi_pool_set_host): code / spec words in the
    layout of text.cuh (base d of a word at bits 2d, A0 C1 G2 T3 = nucl2bit of src/model.c:295-299,
    anything else a separator with code 0), G+C count (src/sequence.c:196-207), separator count,
    zero guard words -- against a numpy restatement, at lengths around the 32-base word and the
    AVX2 block size, with separators at block borders. No GPU involved."""
    L = C.CDLL(str(ROOT / "andi_b200" / "libandi_b200.so"))
    L.andi_host_pack.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.andi_host_pack.restype = None
    rng = np.random.default_rng(11)
    code_of = np.full(256, 0, np.uint64)
    for ch, v in ((b"A", 0), (b"C", 1), (b"G", 2), (b"T", 3)):
        code_of[ch[0]] = v
    for n in (1, 31, 32, 33, 63, 64, 65, 1000, 4097, 100003):
        s = bytearray(np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=n)].tobytes())
        for pos in {0, n - 1, n // 2, 31 % n, 32 % n, 63 % n}:
            if rng.random() < 0.5:
                s[pos] = ord("!")
        s = bytes(s)
        nw = n // 32 + 4
        code, spec = np.full(nw, 0xFFFFFFFFFFFFFFFF, np.uint64), np.full(nw, 0xFFFFFFFFFFFFFFFF, np.uint64)
        gc, sep = C.c_uint64(0), C.c_uint64(0)
        L.andi_host_pack(s, n, code.ctypes.data, spec.ctypes.data, nw, C.byref(gc), C.byref(sep))
        b = np.frombuffer(s, np.uint8)
        nuc = np.isin(b, np.frombuffer(b"ACGT", np.uint8))
        padded_c, padded_s = np.zeros(nw * 32, np.uint64), np.zeros(nw * 32, np.uint64)
        padded_c[:n] = np.where(nuc, code_of[b], 0)
        padded_s[:n] = (~nuc).astype(np.uint64)
        shifts = (np.arange(32, dtype=np.uint64) * np.uint64(2))[None, :]
        want_c = (padded_c.reshape(nw, 32) << shifts).sum(axis=1, dtype=np.uint64)
        want_s = (padded_s.reshape(nw, 32) << shifts).sum(axis=1, dtype=np.uint64)
        assert np.array_equal(code, want_c) and np.array_equal(spec, want_s), n
        assert gc.value == int((nuc & ((b == ord("C")) | (b == ord("G")))).sum()) and sep.value == int((~nuc).sum()), n
