"""Worker of tests/test_gpu_multi.py (launched with torch.distributed.run, one rank per GPU): the
whole matrix of a small pool the way bench.py does it on several ranks -- rank 0 uploads and packs
the pool, its packed planes are broadcast with NCCL, subjects come from a shared queue, rows are
summed to rank 0 -- written to the path given on the command line."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from andi_b200 import driver, native, synth  # noqa: E402

out_path, model = sys.argv[1], sys.argv[2]
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
device = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=device)
store = dist.TCPStore(os.environ["MASTER_ADDR"], int(os.environ["MASTER_PORT"]) + 1, world, is_master=(rank == 0))

seqs = synth.star_phylogeny(11, 40000, [0.0, 0.004, 0.01, 0.02, 0.03, 0.05, 0.08, 0.001, 0.015, 0.025, 0.06], seed=4)
seqs[3] = synth.join_contigs(seqs[3], 4, seed=1)  # one joined genome: the spec plane travels too
n = len(seqs)
ctx = native.Context(local, torch.cuda.current_stream().cuda_stream)
if rank == 0:
    ctx.set_pool(seqs)  # only rank 0 ever sees the characters
received = driver.broadcast_pool(ctx, dist, device, rank)
assert (received > 0) == (rank != 0)
full = torch.zeros((n, n, 17), dtype=torch.int32, device=device)
mine = driver.dynamic_rows(ctx, n, full.data_ptr(), lambda b: store.add("next", b) - b, 2, 0.025, model)
dist.reduce(full, dst=0, op=dist.ReduceOp.SUM)
counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
dist.all_gather(counts, torch.tensor([mine], dtype=torch.int64, device=device))
if rank == 0:
    assert sum(int(c.item()) for c in counts) == n
    np.save(out_path, full.cpu().numpy().view(np.uint32))
dist.barrier()
ctx.close()
dist.destroy_process_group()
