#!/bin/bash
# the headline line once more, with the traffic capture of these sources in place
python bench.py --steps 4 --warmup 3 2> gpurun_out/r2_bench_n1.err | grep '^{' > gpurun_out/r2_bench_n1b.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_n1b.json"))
print("n1", round(d["value"]), round(d["ms_per_step"], 2), round(d["e2e"]["value"]), (d["roofline"]["launch_ms"], d["roofline"]["traffic"], round(d["roofline"]["frac"], 4)), d.get("parity"), d["full_matrix"]["seconds"], d["full_matrix"]["blake2b_of_matrix"])
PY
