import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


# the two 200-nt ESA fixtures of the reference (test/test_esa.c:53-62, 78-88)
ESA_FIXTURE_1 = (
    b"TACGAGCACTGGTGGAATTGATGTC" b"CAGTCTTATATGGCGCACCAGGCTG" b"ATAGTAGTAGCAGTTTGCTTATCTC"
    b"ATCGCGTGTTTCCGGATGACAGAGA" b"TACGTGCACTGGTGGGATTGATGTC" b"TAGTATTATATGGCGCACCAGGATG"
    b"ATAGTAGTAGCAGTTTGCTTATCCC" b"ATCGCGTGTTTGCGGATGACCGAGA"
)
ESA_FIXTURE_2 = (
    b"TACGAGCACTGGTGGAATTGATGTC" b"CAGTCTTATATGGCGCACCAGGCTG" b"ATAGTAGTAGCAGTTTGCTTATCTC"
    b"ATCGCGTGTTTCCGGATGACAGAGA" b"!" b"TACGTGCACTGGTGGGATTGATGTC" b"TAGTATTATATGGCGCACCAGGATG"
    b"ATAGTAGTAGCAGTTTGCTTATCCC" b"ATCGCGTGTTTGCGGATGACCGAGA"
)


@pytest.fixture(scope="session")
def esa_fixtures():
    return [ESA_FIXTURE_1, ESA_FIXTURE_2]


def near_identical_sequences():
    """Outbreak isolates: three and six SNPs in 60 kbp, anchors of tens of kilobases."""
    from andi_b200 import synth

    near = synth.base_genome(60000, seed=31)
    return [synth.ACGT[synth.mutate(near, p, seed=60 + k)].tobytes() for k, p in enumerate((0.0, 5e-5, 1e-4))]


def stress_sequences():
    """Small inputs that exercise the corners the reference's tests and SURVEY 7.3 name:
    substitutions only, indels (diagonal changes), join mode ('!'), repeats, identical and
    unrelated sequences, very short sequences."""
    from andi_b200 import synth

    base = synth.star_phylogeny(4, 30000, [0.0, 0.01, 0.03, 0.08], seed=11)
    out = {"subst": base}
    out["indel"] = [base[0], synth.with_indels(base[1], 40, 30, seed=5), synth.with_indels(base[2], 10, 400, seed=6)]
    out["join"] = [synth.join_contigs(base[0], 7, seed=1), synth.join_contigs(base[1], 3, seed=2), base[2]]
    rep = base[0][:5000] * 3 + base[0][5000:12000] + base[0][2000:4000] + base[0][12000:]
    out["repeat"] = [rep, synth.star_phylogeny(2, len(rep), [0.0, 0.02], seed=3)[1], base[1]]
    out["identical"] = [base[0], base[0], base[1]]
    # realistic repeats: one element in 12 copies (directory buckets beyond the lean scan) and one 3 kbp
    # segment in two copies (tag-2 entries whose candidates both run past a window); the genomes are
    # mutated independently afterwards, so the copies of a subject differ from the query's
    import numpy as np

    g = synth.base_genome(40000, seed=21)
    for i in range(12):
        g[3000 + 2900 * i : 3700 + 2900 * i] = g[1000:1700]
    g[36500:39500] = g[500:3500]
    for i in range(24):  # ... and a short one in 24 copies: buckets beyond what one thread sorts (index build)
        g[2200 + 1450 * i : 2500 + 1450 * i] = g[100:400]
    out["copies"] = [synth.ACGT[synth.mutate(g, p, seed=40 + k)].tobytes() for k, p in enumerate((0.0, 0.01, 0.03))]
    out["near"] = near_identical_sequences()
    unrelated = synth.star_phylogeny(1, 20000, [0.0], seed=99)[0]
    out["unrelated"] = [base[0], unrelated]
    out["short"] = [base[0][:50], base[1][:50], base[0][:300], base[1][10:400]]
    out["lowent"] = [b"ACACACACACAC" * 40 + b"!" + b"ACACACAC" * 30, b"ACACACACACGC" * 40, b"A" * 300 + b"C" + b"A" * 200]
    # revcomp relationship: query is the reverse complement of (a mutated) subject
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    out["revcomp"] = [base[0], base[1].translate(comp)[::-1]]
    return out


@pytest.fixture(scope="session", autouse=True)
def _built_artifacts():
    """The tests need the in-tree libraries; build them if the checkout is fresh."""
    needed = [ROOT / "andi_b200" / "libandi_b200.so", ROOT / "andi_b200" / "libandi_host.so", ROOT / "andi_b200" / "andi"]
    if not all(p.exists() for p in needed):
        import __graft_entry__

        __graft_entry__.build()
