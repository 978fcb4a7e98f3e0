#!/bin/bash
# final 1-GPU measurement job of round 2 (part 2)
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "host_side or esa_arrays or pathological or many_contigs" > gpurun_out/r2_tests_b.log 2>&1; tail -3 gpurun_out/r2_tests_b.log
for pc in 1 2; do ANDI_B200_PART_CTAS=$pc python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu --no-e2e --no-full --rows 16 2>/dev/null | grep '^{' > gpurun_out/r2_bench_c5_ctas$pc.json; done
ANDI_B200_DEPTH_BIAS=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_c4.csv python tools/launch_list.py 64 2100000 3 > /dev/null 2>&1
python bench.py --steps 4 --warmup 3 2> gpurun_out/r2_bench_n1.err | grep '^{' > gpurun_out/r2_bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/r2_bench_ref.err | grep '^{' > gpurun_out/r2_bench_reference_arm.json
python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu --no-e2e --no-full --rows 29 2>/dev/null | grep '^{' > gpurun_out/r2_bench_c2.json
python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu --no-e2e --no-full --rows 109 2>/dev/null | grep '^{' > gpurun_out/r2_bench_c3_kimura.json
python bench.py --workload c3 --model LOGDET --steps 3 --warmup 3 --no-cpu --no-e2e --no-full --rows 109 2>/dev/null | grep '^{' > gpurun_out/r2_bench_c3_logdet.json
python - <<'PY'
import json
for f in ("r2_bench_n1", "r2_bench_reference_arm", "r2_bench_c2", "r2_bench_c3_kimura", "r2_bench_c3_logdet", "r2_bench_c5_ctas1", "r2_bench_c5_ctas2"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"]), round(d["ms_per_step"], 2), d.get("e2e") and round(d["e2e"]["value"]), d.get("roofline") and d["roofline"]["launch_ms"], d.get("esa_build") and d["esa_build"]["ms_per_subject"], d.get("parity"), d.get("full_matrix") and d["full_matrix"]["seconds"])
    except Exception as e:
        print(f, "ERR", e)
PY
