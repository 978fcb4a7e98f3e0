"""GPU box: where does the end-to-end step go? Times andi_pool_set_host and andi_dist_rows (host output) separately."""
import ctypes as C
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import bench
from andi_b200 import native

g, ln, lo, hi, seed, model = bench.WORKLOADS["c4"]
device = torch.device("cuda", 0)
chars, offsets, lens, d = bench.make_pool_device(g, ln, lo, hi, seed, device)
host_pool = torch.empty(chars.numel(), dtype=torch.uint8, pin_memory=True)
host_pool.copy_(chars)
torch.cuda.synchronize()
L = native.load()
import os
other = None
if os.environ.get("E2E_OTHER_CTX"):  # like bench.py: another context with its own copy of the pool is alive
    other = native.Context(0, torch.cuda.current_stream().cuda_stream)
    other.set_pool_device(chars.data_ptr(), offsets, lens)
    o = torch.empty((99, g, 17), dtype=torch.int32, device=device)
    other.dist_rows_device(o.data_ptr(), 0, 99, 0.025, model)
ctx = native.Context(0, torch.cuda.current_stream().cuda_stream) if os.environ.get("E2E_TORCH_STREAM") else native.Context(0)
ptrs = (C.c_char_p * g)(*[C.c_char_p(host_pool.data_ptr() + o) for o in offsets])
lens_c = (C.c_size_t * g)(*lens)
out_host = torch.empty((99, g, 17), dtype=torch.int32, pin_memory=True)
for rep in range(4):
    t0 = time.perf_counter()
    ctx._ck(L.andi_pool_set_host(ctx.h, ptrs, lens_c, g))
    ctx.n = g
    t1 = time.perf_counter()
    ctx._ck(L.andi_dist_rows(ctx.h, 0, 99, 0.025, 1, 0, C.c_void_p(out_host.data_ptr())))
    t2 = time.perf_counter()
    print(f"rep {rep}: pool_set_host {1e3 * (t1 - t0):7.1f} ms   dist_rows(99 rows -> host) {1e3 * (t2 - t1):7.1f} ms", flush=True)
ctx.close()
