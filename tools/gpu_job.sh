#!/bin/bash
# scaling: the bench the way the driver launches it, N ranks on one box
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 4 --warmup 3 2> gpurun_out/r2_bench_n$N.err | grep '^{' > gpurun_out/r2_bench_n$N.json
tail -3 gpurun_out/r2_bench_n$N.err | cut -c1-300
python - $N <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/r2_bench_n{sys.argv[1]}.json"))
print("N", d["n_gpus"], "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 1), "e2e", d["e2e"], "full", d["full_matrix"], "clocks", d["clocks"])
PY
