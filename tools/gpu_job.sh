#!/bin/bash
# N GPUs of one box: the bench under torch.distributed.run (N = number of visible devices)
N=$(nvidia-smi -L | wc -l)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 4 --warmup 3 2> gpurun_out/r2_bench_n$N.err | grep '^{' > gpurun_out/r2_bench_n$N.json
python - $N <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/r2_bench_n{sys.argv[1]}.json"))
print("n", sys.argv[1], round(d["value"]), round(d["ms_per_step"], 2), round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 1), d["full_matrix"]["seconds"], d["full_matrix"]["blake2b_of_matrix"], d["full_matrix"]["rows_per_rank"], d["clocks"])
PY
