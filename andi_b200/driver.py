"""calculate_distances (src/process.c:230-270) for one or several GPUs, one process per GPU.

The matrix rows are independent (src/dist_hack.h:47-72): subject i needs only its own index
and the read-only pool. With W ranks (torch.distributed) the pool is uploaded and packed ONCE, on
rank 0, and its packed planes are broadcast (NCCL over NVLink: 2 bits per base instead of one
byte per base over every rank's PCIe link); subjects are handed out in small batches from a
shared counter (their cost varies with length and divergence) or as static blocks; the rows
are collected on rank 0. There is no exchange step inside the path: the only collectives are
the pool broadcast and the row collection (gloo in the CPU tests of this logic).
(Inside ONE process the same is done by andi_dist_matrix_multi of the C library with host
threads and peer copies; that is what the andi command line uses.)
"""
from __future__ import annotations

import numpy as np


def shard_subjects(n: int, world: int, rank: int):
    """Contiguous, balanced row blocks: rank r owns [begin, end)."""
    base, extra = divmod(n, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def gather_rows(local_rows, n: int, world: int, rank: int, dist=None, device=None):
    """All-gather the (rows_r, n, 17) uint32 blocks of every rank into the (n, n, 17) matrix.
    local_rows may be a numpy array (gloo/CPU) or a torch tensor on `device` (NCCL)."""
    import torch

    t = torch.as_tensor(local_rows) if not isinstance(local_rows, torch.Tensor) else local_rows
    t = t.view(torch.int32) if t.dtype != torch.int32 else t
    if world == 1:
        return t
    sizes = [shard_subjects(n, world, r) for r in range(world)]
    max_rows = max(e - b for b, e in sizes)
    pad = torch.zeros((max_rows, n, 17), dtype=torch.int32, device=t.device)
    pad[: t.shape[0]] = t.reshape(-1, n, 17)
    out = torch.empty((world * max_rows, n, 17), dtype=torch.int32, device=t.device)
    dist.all_gather_into_tensor(out, pad)
    blocks = [out[r * max_rows : r * max_rows + (e - b)] for r, (b, e) in enumerate(sizes)]
    return torch.cat(blocks, dim=0)


def calculate_rows(ctx, n: int, world: int = 1, rank: int = 0, p_value: float = 0.025, model: str = "JC",
                   low_memory: bool = False) -> np.ndarray:
    """This rank's rows of M (src/dist_hack.h:46-72) as a uint32 array (rows, n, 17)."""
    begin, end = shard_subjects(n, world, rank)
    if begin == end:
        return np.empty((0, n, 17), np.uint32)
    return ctx.dist_rows(begin, end, p_value=p_value, model=model, low_memory=low_memory)


class DeviceArray:
    """A raw device allocation as something torch can view (CUDA array interface, read/write)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes // 8,), "typestr": "<i8", "data": (ptr, False), "version": 3}


def broadcast_pool(ctx, dist, device, rank: int, src: int = 0):
    """After rank `src` has set its pool (andi_pool_set_host / _device): every other rank adopts
    the same packed pool through an NCCL broadcast of the planes and of the per-sequence facts
    (lengths, GC fractions -- the doubles the thresholds are computed from --, separator flags).
    Returns the number of plane bytes this rank received."""
    import torch

    head = torch.zeros(3, dtype=torch.int64, device=device)
    view = None
    if rank == src:
        view = ctx.pool_export()
        head = torch.tensor([view["n"], view["words"], int(view["any_separator"])], dtype=torch.int64, device=device)
    dist.broadcast(head, src=src)
    n, words, any_sep = (int(x) for x in head.tolist())
    facts = torch.zeros((3, n), dtype=torch.float64, device=device)
    if rank == src:
        facts[0] = torch.as_tensor(view["lens"].astype(np.float64))  # exact below 2^53
        facts[1] = torch.as_tensor(view["gc"])
        facts[2] = torch.as_tensor(view["has_separator"].astype(np.float64))
    dist.broadcast(facts, src=src)
    planes = []
    for name in ("d_code", "d_spec") if any_sep else ("d_code",):
        if rank == src:
            t = torch.as_tensor(DeviceArray(view[name], words * 8), device=device)
        else:
            t = torch.empty(words, dtype=torch.int64, device=device)
        dist.broadcast(t, src=src)
        planes.append(t)
    if rank == src:
        return 0
    f = facts.cpu().numpy()
    ctx.pool_import({"d_code": planes[0].data_ptr(), "d_spec": planes[1].data_ptr() if any_sep else None, "words": words, "n": n,
                     "lens": f[0].astype(np.uint64), "gc": f[1], "has_separator": f[2].astype(np.int32), "any_separator": any_sep})
    return words * 8 * len(planes)


def dynamic_rows(ctx, n: int, out_dev_ptr: int, take, batch: int, p_value: float = 0.025, model: str = "JC", limit: int | None = None):
    """Rows of M by a shared queue: `take(batch)` returns the first subject of the next batch (an
    atomic fetch-and-add shared by all ranks, e.g. TCPStore.add); rows [b, b + batch) go to
    out_dev_ptr + b * n * 68 (a full n x n matrix on this rank's device). `limit`: end of the
    subjects this queue hands out (default n; a static block passes its own end). Returns the
    number of rows computed."""
    end = n if limit is None else limit
    done = 0
    while True:
        b = take(batch)
        if b >= end:
            return done
        e = min(end, b + batch)
        ctx.dist_rows_device(out_dev_ptr + b * n * 68, b, e, p_value, model)
        done += e - b
