#!/bin/bash
python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2k_gpu_tests.log 2>&1
tail -15 gpurun_out/r2k_gpu_tests.log | cut -c1-200
for m in KIMURA LOGDET; do python bench.py --workload c3 --model $m --steps 3 --warmup 3 --no-cpu --no-e2e --no-full --rows 109 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('c3 $m', d['value'], d['ms_per_step'], d['esa_build']['ms_per_subject'], d['roofline']['launch_ms'], d['cub_calls'])"; done
ANDI_B200_WALK=pipeline python bench.py --workload c3 --model LOGDET --steps 3 --warmup 3 --no-cpu --no-e2e --no-full --rows 109 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('c3 LOGDET pipeline kernel', d['value'], d['ms_per_step'])"
python bench.py --model LOGDET --steps 2 --warmup 2 --no-cpu --no-e2e --no-full 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('c4 LOGDET', d['value'], d['roofline']['launch_ms'])"
