#!/bin/bash
python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py tests/test_gpu_compat.py -m gpu -x -q > gpurun_out/r2_tests_d.log 2>&1; tail -3 gpurun_out/r2_tests_d.log
python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu --no-e2e --no-full --rows 16 2>/dev/null | grep '^{' > gpurun_out/r2_bench_c5.json
python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu --no-e2e --no-full --rows 29 2>/dev/null | grep '^{' > gpurun_out/r2_bench_c2.json
python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu --no-e2e --no-full --rows 109 2>/dev/null | grep '^{' > gpurun_out/r2_bench_c3_kimura.json
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-full 2>/dev/null | grep '^{' > gpurun_out/r2_bench_c4_quick.json
python - <<'PY'
import json
for f in ("r2_bench_c4_quick", "r2_bench_c2", "r2_bench_c3_kimura", "r2_bench_c5"):
    d = json.load(open(f"gpurun_out/{f}.json"))
    print(f, round(d["value"]), round(d["ms_per_step"], 2), d["roofline"]["launch_ms"], d["esa_build"]["ms_per_subject"], d["cub_calls"])
PY
