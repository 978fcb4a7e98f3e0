"""Oracle against the committed golden vectors (tests/golden/, generated from the unmodified
reference by tests/golden/make_golden.py). Needs neither /root/reference nor oracle/_ref."""
import ctypes as C
import gzip
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

import oracle
from conftest import stress_sequences

G = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def golden():
    return json.loads((G / "golden.json").read_text())


@pytest.fixture(scope="module")
def c1():
    fa = gzip.open(G / "c1.fa.gz").read()
    return fa, [s for _, s in oracle.parse_fasta(fa)]


def test_known_answer_vector_of_the_survey(golden, c1):
    # test_fasta -s 1729 -l 100000 -d 0.01, JC, threshold 12
    fa, seqs = c1
    assert hashlib.md5(fa).hexdigest() == golden["c1_md5"] == "9e2c925dc71da4d05e5e5e48209b313d"
    jc = np.array(golden["c1"]["JC"], dtype=np.uint32)
    assert jc[1].tolist() == [24436, 80, 94, 70, 83, 24432, 77, 85, 78, 87, 24406, 88, 84, 85, 81, 25720, 100000]
    assert jc[2].tolist() == [24442, 83, 79, 84, 80, 24425, 87, 85, 94, 77, 24414, 81, 70, 86, 88, 25725, 100000]
    assert golden["c1_threshold"] == 12


@pytest.mark.parametrize("model", ["RAW", "JC", "KIMURA", "LOGDET", "ANI"])
def test_oracle_c1_models(golden, c1, model):
    _, seqs = c1
    got = oracle.rows(seqs, model).reshape(-1, 17)
    assert np.array_equal(got, np.array(golden["c1"][model], dtype=np.uint32))


def test_oracle_c1_estimates(golden):
    m01 = oracle.Model((C.c_uint32 * 16)(*golden["c1"]["JC"][1][:16]), 100000)
    for k, want in golden["c1_estimates_m01"].items():
        got = oracle.lib().orc_estimate(C.byref(m01), oracle.MODELS[k])
        assert got == pytest.approx(want, rel=1e-12, abs=0.0)
    # SURVEY.md 8c quotes these two
    assert golden["c1_estimates_m01"]["RAW"] == pytest.approx(0.0099213889944592248, rel=1e-15)
    assert golden["c1_estimates_m01"]["JC"] == pytest.approx(0.0099875961642709541, rel=1e-15)


def test_oracle_esa_fixtures(golden):
    for ent in golden["esa"]:
        o = oracle.OracleEsa(ent["seq"].encode())
        assert o.rs.decode() == ent["rs"]
        assert o.array("SA").tolist() == ent["SA"]
        assert o.array("LCP").tolist() == ent["LCP"]
        assert o.array("FVC").tolist() == ent["FVC"]
        assert o.array("CLD")[:-1].tolist() == ent["CLD"]
        assert hashlib.sha256(o.array("cache").tobytes()).hexdigest() == ent["cache_sha256"]
        for q, u, c in zip(ent["queries"], ent["get_match"], ent["get_match_cached"]):
            qb = q.encode()
            assert list(o.get_match(qb, "cld")[:3]) == u
            assert list(o.get_match(qb, "cached")[:3]) == c
            assert list(o.get_match(qb, "spec")[:3]) == u
        o.close()


def test_full_depth_sweep(golden):
    # test/test_esa.c:172-203 at its real depth (4^11 queries), both fixtures
    for ent in golden["esa"]:
        o = oracle.OracleEsa(ent["seq"].encode())
        assert oracle.lib().orc_sweep_check(C.byref(o.E), 11) == 0
        o.close()


def test_oracle_stress_sets(golden):
    for name, ss in stress_sequences().items():
        ent = golden["stress"][name]
        assert hashlib.sha256(b"\n".join(ss)).hexdigest() == ent["sha256"], "stress generator drifted: regenerate goldens"
        for model in ("JC", "LOGDET"):
            got = oracle.rows(ss, model).reshape(-1, 17)
            assert np.array_equal(got, np.array(ent[model], dtype=np.uint32)), (name, model)
