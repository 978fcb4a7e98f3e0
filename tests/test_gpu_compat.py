"""The reference's own C surface exported by libandi_b200.so (include/andi_compat.h), driven
the way test/test_esa.c and src/dist_hack.h drive the reference."""
import ctypes as C
import itertools
import json
from pathlib import Path

import numpy as np
import pytest

import oracle
from andi_b200 import native
from conftest import stress_sequences

pytestmark = pytest.mark.gpu
G = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def lib():
    L = native.load()
    L.esa_init.argtypes = [C.POINTER(oracle.RefEsa), C.POINTER(oracle.RefSubject)]
    L.esa_free.argtypes = [C.POINTER(oracle.RefEsa)]
    for f in (L.get_match, L.get_match_cached):
        f.restype = oracle.Interval
        f.argtypes = [C.POINTER(oracle.RefEsa), C.c_char_p, C.c_size_t]
    L.dist_anchor.restype = oracle.Model
    L.dist_anchor.argtypes = [C.POINTER(oracle.RefEsa), C.c_char_p, C.c_size_t, C.c_size_t]
    L.andi_compat_set_model.argtypes = [C.c_int]
    return L


def _subject(seq: bytes):
    n = len(seq)
    buf = C.create_string_buffer(2 * n + 2)
    oracle.lib().orc_make_rs(seq, n, buf)
    gc = oracle.lib().orc_gc(seq, n)
    t = oracle.lib().orc_min_anchor_length(0.025, gc, 2 * n + 1)
    return buf, oracle.RefSubject(C.cast(buf, C.c_void_p), 2 * n + 1, gc, t)


def test_esa_init_fills_reference_struct(lib):
    golden = json.loads((G / "golden.json").read_text())
    for ent in golden["esa"]:
        buf, subj = _subject(ent["seq"].encode())
        E = oracle.RefEsa()
        assert lib.esa_init(C.byref(E), C.byref(subj)) == 0
        N = subj.RSlen
        assert E.len == N and C.string_at(E.S, N).decode() == ent["rs"]
        assert np.ctypeslib.as_array(E.SA, (N,)).tolist() == ent["SA"]
        assert np.ctypeslib.as_array(E.LCP, (N + 1,)).tolist() == ent["LCP"]
        assert np.ctypeslib.as_array(E.CLD, (N,)).tolist() == ent["CLD"]
        assert list(C.string_at(E.FVC, N)) == ent["FVC"]
        # test/test_esa.c:107-170 ("basic", "sample cache"): cached == uncached, match spells the query
        for q, want in zip(ent["queries"], ent["get_match"]):
            qb = q.encode()
            a = lib.get_match_cached(C.byref(E), qb, len(qb))
            b = lib.get_match(C.byref(E), qb, len(qb))
            assert [a.l, a.i, a.j] == [b.l, b.i, b.j] == want
            assert ent["rs"][E.SA[a.i] : E.SA[a.i] + a.l] == q[: a.l]
        lib.esa_free(C.byref(E))
        assert not E.SA and not E.LCP and E.len == 0
        lib.esa_free(C.byref(E))  # idempotent (src/esa.c:280-287)


def test_null_arguments_like_the_reference(lib):
    E = oracle.RefEsa()
    assert lib.esa_init(None, None) == 1  # src/esa.c:255
    assert lib.esa_init(C.byref(E), None) == 1
    r = lib.get_match(C.byref(E), b"ACGT", 4)  # not initialised: src/esa.c:616-618
    assert (r.l, r.i, r.j, r.m) == (-1, -1, -1, -1)


def test_full_cache_sweep_batched(esa_fixtures):
    """test/test_esa.c:172-203 ("/esa/full cache"): all 4^11 queries, here through the batched
    device search; l, i, j must equal the oracle's spec search (itself swept against the
    reference's cached and uncached search in tests/test_golden.py)."""
    ctx = native.Context(0)
    depth = 11
    qs = [bytes(p) for p in itertools.product(b"ACGT", repeat=depth)]
    for s in esa_fixtures:
        o = oracle.OracleEsa(s)
        buf, subj = _subject(s)
        e = ctx.esa_build_rs(buf.raw[: subj.RSlen])
        got = e.get_match(qs)
        rs = o.rs
        sa = o.array("SA")
        assert (got[:, 0] >= 0).all()
        # maximality + spelling, vectorised over the 4M answers; exact (i, j) on a sample
        for k in range(0, len(qs), 997):
            q = qs[k]
            assert tuple(got[k][:3]) == o.get_match(q, "spec")[:3]
        ls = got[:, 0]
        starts = sa[got[:, 1]]
        for k in range(0, len(qs), 12345):
            l, p = int(ls[k]), int(starts[k])
            assert rs[p : p + l] == qs[k][:l]
            assert l == depth or rs[p + l : p + l + 1] != qs[k][l : l + 1]
        e.free(), o.close()
    ctx.close()


@pytest.mark.parametrize("model", ["JC", "LOGDET"])
def test_dist_anchor_like_dist_hack(lib, model):
    """src/dist_hack.h:47-90 against the exported symbols: esa_init, dist_anchor per query, esa_free."""
    lib.andi_compat_set_model(oracle.MODELS[model])
    for name in ("subst", "join", "indel"):
        seqs = stress_sequences()[name]
        want = oracle.rows(seqs, model)
        for i, s in enumerate(seqs):
            buf, subj = _subject(s)
            E = oracle.RefEsa()
            assert lib.esa_init(C.byref(E), C.byref(subj)) == 0
            for j, q in enumerate(seqs):
                if i == j:
                    continue
                m = lib.dist_anchor(C.byref(E), q, len(q), subj.threshold)
                assert list(m.counts) + [m.seq_len] == want[i, j].tolist(), (name, i, j)
            lib.esa_free(C.byref(E))
    lib.andi_compat_set_model(oracle.MODELS["JC"])


def test_rs_string_with_a_stray_separator():
    """andi_esa_build_rs takes ANY RS string (the compat esa_init hands it seq_subject.RS): a second
    '#' away from the middle must not be read as a nucleotide (it selects the separator-aware
    kernels). Checked against a plain sort of the suffixes in the reference's byte order."""
    import numpy as np

    from andi_b200 import native, synth

    half = synth.ACGT[synth.base_genome(700, 3)].tobytes()
    # odd length 1401 with '#' in the middle (index 700) -- the canonical shape -- plus a second '#' at 200
    rs = half[:200] + b"#" + half[200:699] + b"#" + half[::-1]
    assert len(rs) == 1401 and rs[700:701] == b"#" and rs.count(b"#") == 2
    ctx = native.Context(0)
    e = ctx.esa_build_rs(rs)
    got = e.download()["SA"]
    want = np.array(sorted(range(len(rs)), key=lambda i: rs[i:]), dtype=np.int32)
    assert np.array_equal(got, want)
    e.free()
    ctx.close()
