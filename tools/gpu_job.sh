#!/bin/bash
# which boundaries fail with repeats; lean bucket scan against the cooperative scan for every bucket
ANDI_B200_DEBUG_BAD=1 python bench.py --genomes 512 --repeats 30 --steps 1 --warmup 0 --rows 3 --no-cpu --no-e2e --no-full 2>&1 | grep -A6 "andi_b200\]" | head -60
for lib in andi_b200/libandi_b200.so andi_b200/variants/libandi_b200_scan0.so; do
for wl in "--genomes 512 --repeats 30" "--genomes 512" ""; do
ANDI_B200_LIB=$PWD/$lib python bench.py $wl --steps 3 --warmup 2 --no-cpu --no-e2e --no-full 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); n=d['config']['rows_per_step_per_gpu']*d['steps']
print('$lib','$wl',round(d['value']),'walk',round(d['kernel_ms_sums_of_timed_steps']['walk']/n,3),'esa',round(d['kernel_ms_sums_of_timed_steps']['esa']/n,3),'alone',d['roofline']['launch_ms'])"
done; done
