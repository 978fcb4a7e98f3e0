#!/bin/bash
# 2 GPUs: multi-device tests, then the bench at N=2
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -6
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2j_bench_n2.json 2> gpurun_out/r2j_bench_n2.err
tail -3 gpurun_out/r2j_bench_n2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2j_bench_n2.json")); print("n2", d["value"], d["ms_per_step"], d["e2e"], d["full_matrix"])
PY
