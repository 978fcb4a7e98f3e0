"""Several GPUs (skipped on a one-GPU box; run with `gpurun --gpus 2`): the matrix computed on
two devices -- by the C library's own host threads (andi_dist_matrix_multi, what `andi --devices`
uses) and by two torch.distributed ranks (NCCL pool broadcast + shared subject queue, what
bench.py does) -- equals the one-GPU matrix and the oracle's, bit for bit."""
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import oracle
from andi_b200 import native, synth
from test_gpu_cli import write_fasta

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _devices():
    import torch

    return torch.cuda.device_count()


needs_two = pytest.mark.skipif(_devices() < 2, reason="needs two GPUs")


def _pool():
    seqs = synth.star_phylogeny(11, 40000, [0.0, 0.004, 0.01, 0.02, 0.03, 0.05, 0.08, 0.001, 0.015, 0.025, 0.06], seed=4)
    seqs[3] = synth.join_contigs(seqs[3], 4, seed=1)
    return seqs


@needs_two
@pytest.mark.parametrize("model", ["JC", "LOGDET"])
def test_c_library_on_two_devices(model):
    seqs = _pool()
    want = oracle.rows(seqs, model)
    got = native.dist_matrix_multi([0, 1], seqs, model=model)
    assert np.array_equal(got, want)
    one = native.dist_matrix_multi([1], seqs, model=model)  # a single device other than 0
    assert np.array_equal(one, want)


@needs_two
def test_two_nccl_ranks_reproduce_the_single_gpu_matrix(tmp_path):
    seqs = _pool()
    ctx = native.Context(0)
    ctx.set_pool(seqs)
    single = ctx.dist_rows(model="KIMURA")
    ctx.close()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = tmp_path / "matrix.npy"
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), str(ROOT / "tests" / "multi_worker.py"), str(out), "KIMURA"],
                       capture_output=True, timeout=600)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    assert np.array_equal(np.load(out), single)
    assert np.array_equal(single, oracle.rows(seqs, "KIMURA"))


@needs_two
@pytest.mark.skipif(not oracle.REF_ANDI.exists(), reason="oracle/_ref/andi not built")
def test_cli_on_two_devices_prints_the_same_matrix(tmp_path):
    seqs = synth.star_phylogeny(7, 50000, [0.0, 0.002, 0.01, 0.03, 0.05, 0.02, 0.07], seed=8)
    write_fasta(tmp_path / "seven.fa", [(f"g{k}", s) for k, s in enumerate(seqs)])
    ours = ROOT / "andi_b200" / "andi"
    ref = subprocess.run([str(oracle.REF_ANDI), "-t", "1", str(tmp_path / "seven.fa")], capture_output=True)
    one = subprocess.run([str(ours), "--device", "0", str(tmp_path / "seven.fa")], capture_output=True)
    two = subprocess.run([str(ours), "--devices", "0-1", "--progress=always", str(tmp_path / "seven.fa")], capture_output=True)
    assert two.returncode == one.returncode == ref.returncode == 0, two.stderr.decode()
    assert two.stdout == one.stdout == ref.stdout
    assert b"Comparing 7 sequences: 100.0% (42/42), done." in two.stderr


@needs_two
def test_deep_directory_on_every_device():
    """20 Mbp genomes (directory depth 13: the two-level counting sort with 64 KB of dynamic shared
    memory, an attribute that has to be set per device): the matrix from two devices equals the
    one-GPU matrix (which test_large_genomes_against_reference pins to the reference)."""
    seqs = synth.star_phylogeny(3, 20_000_000, [0.0, 0.03, 0.01], seed=4242)
    ctx = native.Context(0)
    ctx.set_pool(seqs)
    single = ctx.dist_rows()
    ctx.close()
    assert np.array_equal(native.dist_matrix_multi([0, 1], seqs), single)
    assert np.array_equal(native.dist_matrix_multi([1], seqs), single)
