#!/bin/bash
# scratch job for the GPU box (edited per run)
python tools/debug_build.py 2>&1 | grep -v " ok" | tail -3
python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py -m gpu -x -q -k "not c5 and not very_large" 2>&1 | tail -3
ANDI_B200_DEPTH_BIAS=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2i_launches_c4.csv python tools/launch_list.py 64 2100000 3 > /dev/null 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu --no-full > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2i_bench.json")); print("c4", d["value"], d["ms_per_step"], d["e2e"], d["esa_build"]["ms_per_subject"], d["roofline"]["launch_ms"], d["cub_calls"])
PY
for w in c2 c3; do python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --no-e2e --no-full --rows 29 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w', d['value'], d['ms_per_step'], d['esa_build']['ms_per_subject'], d['roofline']['launch_ms'], d['cub_calls'])"; done
