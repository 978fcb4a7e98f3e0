"""Multi-rank host logic on CPU (gloo, world_size 2): subject sharding and the row gather of
andi_b200/driver.py. The rows themselves come from the oracle here -- this test is about the
plumbing; GPU parity of the rows is tests/test_gpu_parity.py."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from andi_b200 import driver, synth


def test_shard_subjects_partitions_everything():
    for n in (1, 2, 3, 7, 29, 109, 3085):
        for world in (1, 2, 4, 8):
            spans = [driver.shard_subjects(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[r][1] == spans[r + 1][0] for r in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, seqs, want, ok):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = len(seqs)
    b, e = driver.shard_subjects(n, world, rank)
    local = oracle.rows(seqs, "JC", s_begin=b, s_end=e) if e > b else np.empty((0, n, 17), np.uint32)
    full = driver.gather_rows(local.view(np.int32), n, world, rank, dist=dist)
    got = full.numpy().view(np.uint32)
    ok[rank] = int(np.array_equal(got, want))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_reassembles_the_matrix():
    seqs = synth.star_phylogeny(5, 8000, [0.0, 0.01, 0.02, 0.03, 0.05], seed=77)
    want = oracle.rows(seqs, "JC")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ok = mp.get_context("spawn").Array("i", [0, 0])
    mp.spawn(_worker, args=(2, port, seqs, want, ok), nprocs=2, join=True)
    assert list(ok) == [1, 1]


class _FakeCtx:
    """Stands in for native.Context in the queue test: records which rows it was asked for."""

    def __init__(self):
        self.calls = []

    def dist_rows_device(self, ptr, b, e, p_value, model):
        self.calls.append((ptr, b, e))


def _queue_worker(rank, world, port, n, batch, out):
    store = dist.TCPStore("127.0.0.1", port, world, is_master=(rank == 0))
    ctx = _FakeCtx()
    done = driver.dynamic_rows(ctx, n, 1 << 20, lambda b: store.add("next", b) - b, batch)
    assert done == sum(e - b for _, b, e in ctx.calls)
    for ptr, b, e in ctx.calls:
        assert ptr == (1 << 20) + b * n * 68  # row b of the full matrix, 68 bytes per cell
        for r in range(b, e):
            out[r] += 1
    store.add("finished", 1)
    while store.add("finished", 0) < world:  # rank 0 hosts the store: keep it alive until all are done
        pass


def test_dynamic_queue_hands_every_subject_to_exactly_one_rank():
    """bench.py's whole-matrix leg: subjects are taken in batches from a shared counter
    (TCPStore.add is an atomic fetch-and-add across ranks)."""
    n, batch, world = 109, 4, 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = mp.get_context("spawn").Array("i", [0] * n)
    mp.spawn(_queue_worker, args=(world, port, n, batch, out), nprocs=world, join=True)
    assert list(out) == [1] * n
