#!/bin/bash
# 2 GPUs: the multi-device tests (C threads, NCCL ranks, command line) and the bench under torch.distributed.run
python -m pytest tests/test_gpu_multi.py tests/test_gpu_cli.py -m gpu -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 4 --warmup 3 2> gpurun_out/r2_bench_n2.err | grep '^{' > gpurun_out/r2_bench_n2.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_n2.json"))
print("n2", round(d["value"]), round(d["ms_per_step"], 2), round(d["e2e"]["value"]), d["full_matrix"]["seconds"], d["full_matrix"]["blake2b_of_matrix"], d["full_matrix"]["rows_per_rank"], d.get("parity"))
PY
