#!/bin/bash
# Kernel-tuning sweep on the GPU box: the bench (device-timed part only) once per library variant
# built by `make -C andi_b200/csrc variants`; one JSON line per variant in gpurun_out/variants_<tag>.jsonl
tag=${1:-sweep}
out=gpurun_out/variants_${tag}.jsonl
: > $out
for lib in andi_b200/libandi_b200.so andi_b200/variants/libandi_b200_*.so; do
	name=$(basename $lib .so)
	line=$(ANDI_B200_LIB=$PWD/$lib python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --no-full 2>/dev/null | tail -1)
	echo "{\"variant\": \"$name\", \"line\": $line}" >> $out
	python - "$name" "$line" <<'PY'
import json, sys
d = json.loads(sys.argv[2])
n = d["config"]["rows_per_step_per_gpu"] * d["steps"]
print(f"{sys.argv[1]:32s} {d['value']:10.0f} pairs/s  walk {d["kernel_ms_sums_of_timed_steps"]["walk"] / n:6.3f} ms/subject  esa {d["kernel_ms_sums_of_timed_steps"]["esa"] / n:6.3f}")
PY
done
