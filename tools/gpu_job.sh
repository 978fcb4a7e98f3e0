#!/bin/bash
# deep sort / deep LCP for repeats in the index build: parity, then the repeats workloads
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py tests/test_gpu_compat.py -m gpu -x -q 2>&1 | tail -3
for wl in "--genomes 512 --repeats 30" "" "--repeats 30" "--workload c2" "--workload c2 --repeats 30"; do
python bench.py $wl --steps 3 --warmup 2 --no-cpu --no-e2e --no-full 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); n=d['config']['rows_per_step_per_gpu']*d['steps']
print('$wl',round(d['value']),'ms/step',round(d['ms_per_step'],2),'walk',round(d['kernel_ms_sums_of_timed_steps']['walk']/n,3),'esa',round(d['kernel_ms_sums_of_timed_steps']['esa']/n,3),'alone',d['roofline']['launch_ms'], 'esa alone', d['esa_build']['ms_per_subject'], 'rounds', d['esa_build']['sa_rounds_per_subject'], 'cub', d['cub_calls'])"
done
