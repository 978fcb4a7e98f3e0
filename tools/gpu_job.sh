#!/bin/bash
python tools/debug_build.py 2>&1 | grep -v " ok" | tail -3
python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py -m gpu -x -q > gpurun_out/r2_tests_c.log 2>&1; tail -3 gpurun_out/r2_tests_c.log
python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu --no-e2e --no-full --rows 16 2>/dev/null | grep '^{' > gpurun_out/r2_bench_c5.json
python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu --no-e2e --no-full --rows 29 2>/dev/null | grep '^{' > gpurun_out/r2_bench_c2.json
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-full 2>/dev/null | grep '^{' > gpurun_out/r2_bench_c4_quick.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_c5.csv python tools/launch_list.py 2 120000000 1 > /dev/null 2>&1
python - <<'PY'
import json
for f in ("r2_bench_c4_quick", "r2_bench_c2", "r2_bench_c5"):
    d = json.load(open(f"gpurun_out/{f}.json"))
    print(f, round(d["value"]), round(d["ms_per_step"], 2), d["roofline"]["launch_ms"], d["esa_build"]["ms_per_subject"], d["cub_calls"])
PY
