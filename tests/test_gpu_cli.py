"""The andi command line built on the GPU library (andi_b200/andi) against the reference
binary (oracle/_ref/andi): byte-identical stdout and equal exit status on the same files
(test/test_extra.sh, test_join.sh, nan.sh, low_homo.sh restated as exact diffs)."""
import gzip
import subprocess
from pathlib import Path

import pytest

import oracle
from andi_b200 import synth
from conftest import stress_sequences

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not oracle.REF_ANDI.exists(), reason="oracle/_ref/andi not built")]
ROOT = Path(__file__).resolve().parent.parent
OURS = ROOT / "andi_b200" / "andi"
G = ROOT / "tests" / "golden"


def write_fasta(path, records, width=70):
    with open(path, "w") as f:
        for name, seq in records:
            f.write(f">{name} some comment\n")
            s = seq.decode()
            for k in range(0, len(s), width):
                f.write(s[k : k + width] + "\n")


def both(args, stdin=None, cwd=None):
    a = subprocess.run([str(oracle.REF_ANDI), "-t", "1", *args], input=stdin, capture_output=True, cwd=cwd)
    b = subprocess.run([str(OURS), *args], input=stdin, capture_output=True, cwd=cwd)
    return a, b


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    d = tmp_path_factory.mktemp("cli")
    (d / "c1.fa").write_bytes(gzip.open(G / "c1.fa.gz").read())
    seqs = synth.star_phylogeny(5, 60000, [0.0, 0.0005, 0.01, 0.03, 0.08], seed=21)
    write_fasta(d / "five.fa", [(f"genome_number_{k}_long_name", s) for k, s in enumerate(seqs)])
    for k, s in enumerate(seqs[:3]):
        parts = s.split(b"A" * 0) if False else [s[: len(s) // 3], s[len(s) // 3 : len(s) // 2], s[len(s) // 2 :]]
        write_fasta(d / f"asm{k}.part.fasta", [(f"contig{c}", p) for c, p in enumerate(parts)])
    unrelated = stress_sequences()["unrelated"]
    write_fasta(d / "u0.fa", [("U0", unrelated[0])])
    write_fasta(d / "u1.fa", [("U1", unrelated[1])])
    (d / "fof.txt").write_text(f"{d / 'five.fa'}\n\n{d / 'c1.fa'}\n")
    return d


@pytest.mark.parametrize("args", [
    [], ["-m", "RAW"], ["-m", "Kimura"], ["-m", "LOGDET"], ["-m", "ani"], ["-l"], ["-v"], ["-vv"],
    ["--truncate-names"], ["-p", "0.1"], ["-p", "0.001", "-m", "RAW"],
])
def test_same_output_on_five_genomes(files, args):
    a, b = both([*args, str(files / "five.fa")])
    assert b.stdout == a.stdout, b.stderr.decode()
    assert b.returncode == a.returncode


def test_config1_prints_the_known_matrix(files):
    a, b = both([str(files / "c1.fa")])
    assert b.stdout == a.stdout == b"2\nS0         0.0000 0.0100\nS1         0.0100 0.0000\n"


def test_join_mode_and_several_files(files):
    names = [str(files / f"asm{k}.part.fasta") for k in range(3)]
    a, b = both(["-j", "-m", "RAW", *names])
    assert b.stdout == a.stdout and b.returncode == a.returncode
    a, b = both([str(files / "five.fa"), str(files / "c1.fa")])
    assert b.stdout == a.stdout


def test_file_of_filenames_and_stdin(files):
    a, b = both(["--file-of-filenames", str(files / "fof.txt")])
    assert b.stdout == a.stdout and b.returncode == a.returncode
    data = (files / "five.fa").read_bytes()
    a, b = both([], stdin=data)
    assert b.stdout == a.stdout
    a, b = both(["-"], stdin=data)
    assert b.stdout == a.stdout


def test_unrelated_sequences_warn_like_the_reference(files):
    # test/nan.sh / low_homo.sh: nan or "very little homology" + failure exit status
    a, b = both(["-j", str(files / "u0.fa"), str(files / "u1.fa")])
    assert b.stdout == a.stdout and b.returncode == a.returncode
    for word in (b"nan", b"homology"):
        assert (word in a.stderr) == (word in b.stderr)


def test_bootstrap_prints_b_matrices(files):
    b = subprocess.run([str(OURS), "-b", "4", "--seed", "7", str(files / "five.fa")], capture_output=True)
    lines = b.stdout.decode().splitlines()
    assert len(lines) == 4 * 6 and lines[0] == "5" and lines[6] == "5"
    first = [float(x) for x in lines[3].split()[1:]]
    boot = [float(x) for x in lines[9].split()[1:]]
    assert all(abs(x - y) <= 0.15 * max(x, 1e-3) + 2e-4 for x, y in zip(first, boot))  # resampled, not wildly off


def test_errors(files):
    a, b = both([str(files / "u0.fa")])  # fewer than two sequences
    assert a.returncode == b.returncode == 1 and b"nothing to compare" in b.stderr
    a, b = both([str(files / "does_not_exist.fa"), str(files / "c1.fa")])
    assert b.stdout == a.stdout and a.returncode == b.returncode == 1
