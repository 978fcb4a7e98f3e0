"""Bit-exact parity at the SHAPES bench.py measures (BASELINE.json configs 2-5), against the
unmodified reference compiled into oracle/_ref (src/esa.c:254-277, src/process.c:141-214 driven
like src/dist_hack.h:34-72): genome length, divergence range, directory depth and chunk length
are those of the benched runs, only the number of genomes is cut down to what the CPU finishes
in seconds. Skipped where oracle/_ref was not built (it needs /root/reference at build time)."""
import numpy as np
import pytest

import oracle
from andi_b200 import native, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = native.Context(0)
    yield c
    c.close()


def _need_ref():
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built")


def _assert_rows(got, want, what):
    bad = np.argwhere((got != want).any(axis=2))
    assert len(bad) == 0, (what, len(bad), bad[:5].tolist(), got[tuple(bad[0])].tolist(), want[tuple(bad[0])].tolist())


def test_c4_shape(ctx, monkeypatch):
    """Config 4 as benched: 2.1 Mbp genomes, d_k ~ U[0.005, 0.02], JC, directory depth 12 (what
    choose_depth picks once a subject is walked by a gigabase of queries) and the chunk length the
    3085-genome pool gets (8192)."""
    _need_ref()
    monkeypatch.setenv("ANDI_B200_DEPTH_BIAS", "1")
    monkeypatch.setenv("ANDI_B200_CHUNK", "8192")
    n = 8
    seqs = synth.star_phylogeny(n, 2_100_000, synth.config_divergences(n, 0.005, 0.02, 3085), seed=3085)
    want, _ = oracle.ref_rows(seqs, "JC", threads=8)
    ctx.set_pool(seqs)
    _assert_rows(ctx.dist_rows(model="JC"), want, "c4")
    # the round-1 kernel stays selectable and must agree as well
    monkeypatch.setenv("ANDI_B200_WALK", "pipeline")
    _assert_rows(ctx.dist_rows(model="JC"), want, "c4 pipeline kernel")


@pytest.mark.parametrize("model", ["JC", "KIMURA", "LOGDET"])
def test_c2_c3_shape(ctx, model):
    """Configs 2 and 3: 5 Mbp genomes (N = 10 M: the bucketing path for large texts), d_k up to
    5 %, JC / KIMURA / LOGDET, default depth and chunk length."""
    _need_ref()
    n = 4
    seqs = synth.star_phylogeny(n, 5_000_000, [0.01, 0.05, 0.03, 0.02], seed=109)
    want, _ = oracle.ref_rows(seqs, model, threads=4)
    ctx.set_pool(seqs)
    _assert_rows(ctx.dist_rows(model=model), want, model)


def test_c5_shape(ctx):
    """Config 5: 120 Mbp (N = 240 000 001, directory depth 14), one subject against two queries,
    low-memory mode; the reference's suffix array comes from the oracle's divsufsort stand-in."""
    _need_ref()
    n = 120_000_000
    seqs = synth.star_phylogeny(3, n, [0.0, 0.01, 0.05], seed=16)
    want, _ = oracle.ref_rows(seqs, "JC", s_begin=0, s_end=1, threads=2, low_memory=True)
    ctx.set_pool(seqs)
    got = ctx.dist_rows(s_begin=0, s_end=1, low_memory=True)
    _assert_rows(got, want, "c5")
