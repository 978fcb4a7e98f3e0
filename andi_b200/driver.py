"""calculate_distances (src/process.c:230-270) for one or several GPUs.

The matrix rows are independent (src/dist_hack.h:47-72): subject i needs only its own index
and the read-only pool. With W ranks (one process per GPU, torch.distributed) every rank
holds the whole packed pool and computes the rows of its subjects; the row blocks are then
gathered to rank 0. There is no exchange step inside the path, so the only collective is that
gather (NCCL over NVLink on GPUs; gloo in the CPU tests of the sharding logic).
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
