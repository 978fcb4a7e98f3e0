#!/bin/bash
# scratch job for the GPU box (edited per run)
python tools/debug_build.py 2>&1 | grep -v " ok" | tail -5
python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py tests/test_gpu_compat.py -m gpu -x -q 2>&1 | tail -5
ANDI_B200_DEPTH_BIAS=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2g_launches_c4.csv python tools/launch_list.py 64 2100000 3 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2g_launches_c5.csv python tools/launch_list.py 2 120000000 1 > /dev/null 2>&1
python bench.py --workload c5 --steps 2 --warmup 1 --no-cpu --no-e2e --no-full --rows 4 | cut -c1-150
python - <<'PY'
import torch, time
x = torch.empty(6_000_000_000, dtype=torch.uint8, pin_memory=True)
y = torch.empty_like(x, device="cuda")
for _ in range(2):
    torch.cuda.synchronize(); t = time.perf_counter(); y.copy_(x, non_blocking=True); torch.cuda.synchronize()
    print("H2D pinned 6 GB: %.1f GB/s" % (6.0 / (time.perf_counter() - t)))
PY
lscpu | grep -E "Model name|^CPU\(s\)|Flags" | cut -c1-400
