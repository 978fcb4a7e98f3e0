"""INTEGRATION.md route 2 for real: the reference's own program -- src/andi.c, io.c, sequence.c,
model.c and pfasta, compiled from /root/reference WITHOUT src/esa.c and src/process.c -- linked
against libandi_b200.so (`make -C oracle relink` -> oracle/_ref/andi_relinked). The library's
calculate_distances (src/process.h:11) runs the matrix on the GPU and hands it to the program's
own print_distances / print_coverages; stdout must equal the unmodified reference binary's."""
import subprocess
from pathlib import Path

import pytest

import oracle
from andi_b200 import synth
from test_gpu_cli import write_fasta

RELINKED = oracle.HERE / "_ref" / "andi_relinked"
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (RELINKED.exists() and oracle.REF_ANDI.exists()), reason="oracle/_ref not built")]


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    d = tmp_path_factory.mktemp("relink")
    seqs = synth.star_phylogeny(5, 60000, [0.0, 0.0005, 0.01, 0.03, 0.08], seed=21)
    write_fasta(d / "five.fa", [(f"genome_{k}", s) for k, s in enumerate(seqs)])
    for k, s in enumerate(seqs[:3]):
        parts = [s[: len(s) // 3], s[len(s) // 3 : len(s) // 2], s[len(s) // 2 :]]
        write_fasta(d / f"asm{k}.fasta", [(f"contig{c}", p) for c, p in enumerate(parts)])
    return d


@pytest.mark.parametrize("args", [[], ["-m", "RAW"], ["-m", "Kimura"], ["-m", "LogDet"], ["-l"], ["-v"], ["-p", "0.1"],
                                  ["--truncate-names"]])
def test_relinked_reference_prints_the_same_matrix(files, args):
    a = subprocess.run([str(oracle.REF_ANDI), "-t", "1", *args, str(files / "five.fa")], capture_output=True)
    b = subprocess.run([str(RELINKED), "-t", "1", *args, str(files / "five.fa")], capture_output=True)
    assert b.returncode == a.returncode, b.stderr.decode()
    assert b.stdout == a.stdout, b.stderr.decode()


def test_relinked_reference_join_mode_and_progress(files):
    names = [str(files / f"asm{k}.fasta") for k in range(3)]
    a = subprocess.run([str(oracle.REF_ANDI), "-t", "1", "-j", *names], capture_output=True)
    b = subprocess.run([str(RELINKED), "-j", "--progress=always", *names], capture_output=True)
    assert b.returncode == a.returncode, b.stderr.decode()
    assert b.stdout == a.stdout
    # src/dist_hack.h:37-43,74-95: the progress line ends at 100 % of n*n-n comparisons
    assert b"Comparing 3 sequences: 100.0% (6/6), done." in b.stderr
