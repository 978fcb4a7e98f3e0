#!/bin/bash
# final pass, 1 GPU: full -m gpu suite, smoke, headline bench, traffic capture tied to the final sources
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests.log 2>&1; tail -3 gpurun_out/r2_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
tools/capture_traffic.sh r2 > /dev/null
python tools/ncu_summary.py gpurun_out/r2_walk.ncu-rep > gpurun_out/r2_walk_ncu_summary.txt
ncu -i gpurun_out/r2_walk.ncu-rep --page source --csv --print-source sass > gpurun_out/r2_walk_source.csv 2>/dev/null
python tools/ncu_regions.py gpurun_out/r2_walk_source.csv "k_walk_v3<(int)1" > gpurun_out/r2_walk_regions.txt
cp gpurun_out/walk_traffic.json profiles/walk_traffic.json
python bench.py --steps 4 --warmup 3 2> gpurun_out/r2_bench_n1.err | grep '^{' > gpurun_out/r2_bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/r2_bench_ref.err | grep '^{' > gpurun_out/r2_bench_reference_arm.json
python - <<'PY'
import json
for f in ("r2_bench_n1", "r2_bench_reference_arm"):
    d = json.load(open(f"gpurun_out/{f}.json"))
    print(f, round(d["value"]), round(d["ms_per_step"], 2), d.get("e2e") and round(d["e2e"]["value"]), d.get("roofline") and (d["roofline"]["launch_ms"], d["roofline"]["traffic"], d["roofline"]["frac"]), d.get("parity"), d.get("full_matrix") and d["full_matrix"]["seconds"])
PY
