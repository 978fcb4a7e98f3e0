"""Parity of the CUDA path (through the C ABI) with the oracle on the same inputs.
Integer work: every comparison is bit-exact."""
import gzip
import json
import random
from pathlib import Path

import numpy as np
import pytest

import oracle
from andi_b200 import native, synth
from conftest import stress_sequences

pytestmark = pytest.mark.gpu
G = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def ctx():
    c = native.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def golden():
    return json.loads((G / "golden.json").read_text())


def _esa_cases(esa_fixtures):
    groups = stress_sequences()
    cases = list(esa_fixtures)
    for name in ("subst", "join", "repeat", "short", "lowent", "indel"):
        cases.extend(groups[name])
    return cases


def test_esa_arrays_bit_identical(ctx, esa_fixtures):
    """E1-E5: SA, LCP, CLD, FVC and the 4^10 prefix cache equal the oracle's (= the reference's)."""
    cases = _esa_cases(esa_fixtures)
    ctx.set_pool(cases)
    for k, s in enumerate(cases):
        o = oracle.OracleEsa(s)
        e = ctx.esa_build(k, full=True)
        got = e.download(full=True)
        assert np.array_equal(got["SA"], o.array("SA")), ("SA", k)
        assert np.array_equal(got["LCP"], o.array("LCP")), ("LCP", k)
        assert np.array_equal(got["FVC"], o.array("FVC")), ("FVC", k)
        assert np.array_equal(got["CLD"][:-1], o.array("CLD")[:-1]), ("CLD", k)
        assert np.array_equal(got["cache"], o.array("cache")), ("cache", k)
        e.free(), o.close()


def test_esa_from_rs_string_matches_golden(ctx, golden):
    """esa_init's own calling convention (an RS string) against the reference's arrays."""
    for ent in golden["esa"]:
        e = ctx.esa_build_rs(ent["rs"].encode(), full=True)
        got = e.download(full=True)
        assert got["SA"].tolist() == ent["SA"]
        assert got["LCP"].tolist() == ent["LCP"]
        assert got["FVC"].tolist() == ent["FVC"]
        assert got["CLD"][:-1].tolist() == ent["CLD"]
        qs = [q.encode() for q in ent["queries"]]
        m = e.get_match(qs)
        assert m[:, :3].tolist() == ent["get_match"]
        e.free()


def test_get_match_spec(ctx, esa_fixtures):
    """E6: longest prefix match and its exact SA range, against the oracle's spec search."""
    rng = random.Random(7)
    groups = stress_sequences()
    cases = list(esa_fixtures) + groups["join"][:2] + groups["repeat"][:1] + groups["subst"][:1] + groups["lowent"][:1]
    ctx.set_pool(cases)
    for k, s in enumerate(cases):
        o = oracle.OracleEsa(s)
        rs = o.rs
        qs = []
        for _ in range(3000):
            ln = rng.choice([1, 2, 5, 9, 10, 11, 12, 13, 17, 25, 40, 90])
            r = rng.random()
            if r < 0.6:
                p = rng.randrange(0, len(rs) - 1)
                q = bytearray(rs[p : p + ln].replace(b"#", b"A").replace(b";", b"!"))
                if q and rng.random() < 0.5:
                    q[rng.randrange(len(q))] = rng.choice(b"ACGT")
                q = bytes(q)
            else:
                q = bytes(rng.choice(b"ACGT") for _ in range(ln))
            if q:
                qs.append(q)
        e = ctx.esa_build(k)
        got = e.get_match(qs)
        for q, g in zip(qs, got):
            assert tuple(g[:3]) == o.get_match(q, "spec")[:3], (k, q)
        e.free(), o.close()


@pytest.mark.parametrize("model", ["JC", "LOGDET", "RAW", "KIMURA", "ANI"])
def test_rows_match_oracle_on_stress_sets(ctx, model):
    """P1-P4, M1, M2, D1: whole rows of the matrix, every corner case group."""
    for name, seqs in stress_sequences().items():
        ctx.set_pool(seqs)
        got = ctx.dist_rows(model=model)
        want = oracle.rows(seqs, model)
        assert np.array_equal(got, want), (name, model)
        low = ctx.dist_rows(model=model, low_memory=True)  # test/test_extra.sh:19-22
        assert np.array_equal(low, want), (name, model, "low-memory")


def test_config1_known_answer(ctx, golden):
    """BASELINE.json configs[0]: the survey's known-answer vector, from the committed FASTA."""
    fa = gzip.open(G / "c1.fa.gz").read()
    seqs = [s for _, s in oracle.parse_fasta(fa)]
    ctx.set_pool(seqs)
    for model in ("JC", "LOGDET", "RAW"):
        got = ctx.dist_rows(model=model).reshape(-1, 17)
        assert np.array_equal(got, np.array(golden["c1"][model], dtype=np.uint32)), model
    # legacy call shapes: one index, explicit query ids / one host query string
    e = ctx.esa_build(0)
    t = native.threshold(0.025, ctx.pool_info(0)[1], 2 * len(seqs[0]) + 1)
    assert t == golden["c1_threshold"]
    row = ctx.dist_row(e, [1], t, "JC")
    assert row[0].tolist() == golden["c1"]["JC"][1]
    one = ctx.dist_anchor(e, seqs[1], t, "JC")
    assert one.tolist() == golden["c1"]["JC"][1]
    e.free()


def test_p_value_and_explicit_threshold(ctx):
    seqs = stress_sequences()["subst"]
    ctx.set_pool(seqs)
    for p in (0.1, 0.001):
        assert np.array_equal(ctx.dist_rows(p_value=p), oracle.rows(seqs, "JC", p_value=p))
    o = oracle.OracleEsa(seqs[0])
    e = ctx.esa_build(0)
    for t in (5, 9, 14, 30):  # below / above the directory depth
        assert np.array_equal(ctx.dist_anchor(e, seqs[2], t, "JC"), o.dist_anchor(seqs[2], t, "JC", spec=True)), t
    e.free(), o.close()


def test_medium_genomes_with_properties(ctx):
    """1 Mbp genomes: parity with the oracle plus size-independent properties."""
    seqs = synth.star_phylogeny(4, 1_000_000, [0.0, 0.005, 0.02, 0.05], seed=2024)
    ctx.set_pool(seqs)
    got = ctx.dist_rows()
    want = oracle.rows(seqs, "JC", s_begin=0, s_end=2)
    assert np.array_equal(got[:2], want)
    tot = got[:, :, :16].sum(axis=2)
    assert (got[:, :, 16] >= tot)[~np.eye(4, dtype=bool)].all()  # covered <= query length
    # identical genomes: everything is covered, nothing is a substitution
    ident = [seqs[1], seqs[1], seqs[2]]
    ctx.set_pool(ident)
    g2 = ctx.dist_rows(s_begin=0, s_end=1)
    assert g2[0, 1, :16].sum() == len(seqs[1]) and g2[0, 1, [0, 5, 10, 15]].sum() == len(seqs[1])


@pytest.mark.parametrize("chunk", [64, 257, 1000, 4096])
def test_chunk_boundaries_are_exact(ctx, chunk, monkeypatch):
    """The chunked walk must not depend on where the chunk boundaries fall (walk_kernels.cuh):
    odd chunk lengths, chunks shorter than an anchor, chunks longer than the sequences."""
    monkeypatch.setenv("ANDI_B200_CHUNK", str(chunk))
    for name, seqs in stress_sequences().items():
        ctx.set_pool(seqs)
        for model in ("JC", "LOGDET"):
            got = ctx.dist_rows(model=model)
            want = oracle.rows(seqs, model)
            assert np.array_equal(got, want), (name, model, chunk)


@pytest.mark.parametrize("bias", [-2, 1, 2])
def test_directory_depth_does_not_change_results(ctx, bias, monkeypatch):
    """The k-mer directory depth K is a tuning choice (one level deeper for large pools,
    choose_depth in andi_b200.cu): shallower and deeper directories must give the same rows."""
    monkeypatch.setenv("ANDI_B200_DEPTH_BIAS", str(bias))
    for name, seqs in stress_sequences().items():
        ctx.set_pool(seqs)
        for model in ("JC", "ANI"):
            assert np.array_equal(ctx.dist_rows(model=model), oracle.rows(seqs, model)), (name, model, bias)


def test_large_genomes_against_reference(ctx):
    """20 Mbp genomes (N = 40 M): 32-bit index arithmetic, directory depth 13, several chunks
    per thread. Checked against the reference itself (oracle/_ref) because the oracle's
    doubling sorter is too slow at this size; skipped where _ref was not built."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built")
    seqs = synth.star_phylogeny(2, 20_000_000, [0.0, 0.03], seed=4242)
    want, _ = oracle.ref_rows(seqs, "JC", threads=2)
    ctx.set_pool(seqs)
    got = ctx.dist_rows()
    assert np.array_equal(got, want)


def test_very_large_genome_properties(ctx):
    """120 Mbp (config 5 shape, N = 240 000 001): size-independent properties only -- identical
    genomes are fully covered with no substitutions; a copy with k planted substitutions far
    apart reports exactly k mismatches."""
    n = 120_000_000
    base = synth.base_genome(n, 99)
    a = synth.ACGT[base].tobytes()
    rng = np.random.default_rng(5)
    pos = np.sort(rng.choice(n // 1000, size=500, replace=False)) * 1000 + 500
    mut = base.copy()
    mut[pos] = (mut[pos] + 1) & 3
    b = synth.ACGT[mut].tobytes()
    ctx.set_pool([a, a, b])
    got = ctx.dist_rows(s_begin=0, s_end=1)
    ident, sub = got[0, 1], got[0, 2]
    assert ident[:16].sum() == n and ident[[0, 5, 10, 15]].sum() == n
    assert sub[:16].sum() == n and sub[:16].sum() - sub[[0, 5, 10, 15]].sum() == 500


def test_many_contigs(ctx):
    """Join mode with hundreds of contigs: thousands of suffixes start with (or run into) a
    separator and share a padded bucket key; they are ordered through the padded-suffix list
    (sa_bucket.cuh). Includes adjacent separators and duplicated contigs (equal 21-character
    order keys, settled by direct comparison). Arrays and rows against the oracle."""
    rng = np.random.default_rng(77)
    codes = synth.base_genome(60_000, 5)
    base = synth.ACGT[codes].tobytes()
    a = bytearray(synth.join_contigs(base, 400, seed=9))
    for k in rng.choice(len(a) - 2, size=20, replace=False):  # a few adjacent separators
        a[k] = a[k + 1] = ord("!")
    contig = synth.ACGT[synth.base_genome(3000, 6)].tobytes()
    other = synth.ACGT[synth.base_genome(20_000, 7)].tobytes()
    b = b"!".join([contig, other[:7000], contig, contig, other[7000:], contig[:1500]])
    c = synth.join_contigs(synth.ACGT[synth.mutate(codes, 0.02, seed=3)].tobytes(), 250, seed=10)
    seqs = [bytes(a), b, c, base]
    ctx.set_pool(seqs)
    for k, s in enumerate(seqs[:3]):
        o = oracle.OracleEsa(s)
        e = ctx.esa_build(k)
        got = e.download()
        assert np.array_equal(got["SA"], o.array("SA")), ("SA", k)
        assert np.array_equal(got["LCP"], o.array("LCP")), ("LCP", k)
        e.free(), o.close()
    for model in ("JC", "LOGDET"):
        assert np.array_equal(ctx.dist_rows(model=model), oracle.rows(seqs, model)), model


def test_large_joined_genomes_against_reference(ctx):
    """5 Mbp draft-assembly shape (N = 10 M > 8 M: radix bucketing path) with 300 contigs each,
    against the reference itself: suffix array, LCP and the rows."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built")
    plain = synth.star_phylogeny(2, 5_000_000, [0.0, 0.02], seed=515)
    seqs = [synth.join_contigs(plain[0], 300, seed=1), synth.join_contigs(plain[1], 300, seed=2)]
    ctx.set_pool(seqs)
    h = oracle.RefEsaHandle(seqs[0])
    e = ctx.esa_build(0)
    got = e.download()
    assert np.array_equal(got["SA"], h.array("SA"))
    assert np.array_equal(got["LCP"], h.array("LCP"))
    e.free()
    want, _ = oracle.ref_rows(seqs, "JC", threads=2)
    assert np.array_equal(ctx.dist_rows(), want)


def test_pathological_repeats(ctx):
    """Low-complexity and tandem-repeat texts: every suffix ties for thousands of characters, so
    the bucket sorter hands everything to the doubling rounds and the direct LCP overflows into
    the phi/Kasai path (index_host.cuh). Arrays and rows must still equal the oracle's."""
    rng = np.random.default_rng(8)
    poly = bytearray(b"A" * 60000 + b"C" + b"A" * 60000)
    tandem = bytearray(b"ACGT" * 30000)
    for k in rng.choice(len(tandem), size=25, replace=False):
        tandem[k] = ord("G") if tandem[k] != ord("G") else ord("T")
    mixed = bytes(synth.ACGT[synth.base_genome(40000, 5)].tobytes()) + bytes(tandem[:40000]) + b"T" * 3000
    seqs = [bytes(poly), bytes(tandem), mixed, bytes(poly[:50000]) + b"!" + bytes(tandem[:30000])]
    ctx.set_pool(seqs)
    for k in (0, 1, 3):
        o = oracle.OracleEsa(seqs[k])
        e = ctx.esa_build(k, full=True)
        got = e.download(full=True)
        assert np.array_equal(got["SA"], o.array("SA")), ("SA", k)
        assert np.array_equal(got["LCP"], o.array("LCP")), ("LCP", k)
        assert np.array_equal(got["CLD"][:-1], o.array("CLD")[:-1]), ("CLD", k)
        assert np.array_equal(got["FVC"], o.array("FVC")), ("FVC", k)
        assert np.array_equal(got["cache"], o.array("cache")), ("cache", k)
        e.free(), o.close()
    for model in ("JC", "LOGDET"):
        assert np.array_equal(ctx.dist_rows(model=model), oracle.rows(seqs, model)), model


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_randomised_pools(ctx, seed, monkeypatch):
    """Many small random inputs in one pool: random lengths (30 .. 6000), divergences, indels,
    contig separators, reverse complements, duplicates -- all ordered pairs against the oracle,
    with a chunk length small enough that most pairs span several chunks."""
    monkeypatch.setenv("ANDI_B200_CHUNK", "512")
    rng = np.random.default_rng(seed)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    base = synth.ACGT[synth.base_genome(6000, 100 + seed)].tobytes()
    seqs = []
    for k in range(26):
        ln = int(rng.integers(30, 6000))
        start = int(rng.integers(0, 6000 - ln + 1))
        codes = np.array([b"ACGT".index(c) for c in base[start : start + ln]], np.uint8)
        s = synth.ACGT[synth.mutate(codes, float(rng.uniform(0, 0.08)), seed * 1000 + k)].tobytes()
        r = rng.random()
        if r < 0.25:
            s = synth.with_indels(s, int(rng.integers(1, 6)), 20, seed * 77 + k)
        elif r < 0.45 and len(s) > 200:
            s = synth.join_contigs(s, int(rng.integers(2, 6)), seed * 55 + k)
        elif r < 0.6:
            s = s.translate(comp)[::-1]
        elif r < 0.65 and seqs:
            s = seqs[int(rng.integers(0, len(seqs)))]
        seqs.append(s)
    for model in ("JC", "LOGDET"):
        ctx.set_pool(seqs)
        got = ctx.dist_rows(model=model)
        want = oracle.rows(seqs, model)
        bad = np.argwhere((got != want).any(axis=2))
        assert len(bad) == 0, (model, bad[:5].tolist(), [len(seqs[i]) for i in bad[0]])


def test_host_side_packing_gives_the_same_pool(ctx, monkeypatch):
    """andi_pool_set_host can pack the pool on the host (host_pack.c: AVX2 / scalar, OpenMP) before
    the upload (ANDI_B200_HOST_PACK=1) instead of on the device: same GC fractions, separator flags
    and rows, on inputs with separators, odd lengths and lengths around the 32-base word size."""
    groups = stress_sequences()
    seqs = groups["join"] + groups["short"] + [groups["subst"][1][:31], groups["subst"][1][:32], groups["subst"][1][:33],
                                               groups["subst"][2][:64] + b"!" + groups["subst"][2][64:131]]
    ctx.set_pool(seqs)
    device_info = [ctx.pool_info(k) for k in range(len(seqs))]
    device_rows = ctx.dist_rows(model="JC")
    monkeypatch.setenv("ANDI_B200_HOST_PACK", "1")
    ctx.set_pool(seqs)
    assert [ctx.pool_info(k) for k in range(len(seqs))] == device_info
    assert np.array_equal(ctx.dist_rows(model="JC"), device_rows)
    assert np.array_equal(device_rows, oracle.rows(seqs, "JC"))
