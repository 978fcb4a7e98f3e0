#!/bin/bash
# final pass, 1 GPU (about 9 GPU-minutes): full -m gpu suite, smoke, traffic capture tied to the sources, headline line +
# reference arm, the other configurations. `gpurun --timeout 2400 -- tools/gpu_job.sh`; results land in gpurun_out/,
# copy what is to be kept into profiles/ (profiles/README.md).
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests.log 2>&1; tail -3 gpurun_out/r2_gpu_tests.log
grep -q " passed" gpurun_out/r2_gpu_tests.log && ! grep -q "failed" gpurun_out/r2_gpu_tests.log || exit 1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
tools/capture_traffic.sh r2 > /dev/null
python tools/ncu_summary.py gpurun_out/r2_walk.ncu-rep > gpurun_out/r2_walk_ncu_summary.txt
ncu -i gpurun_out/r2_walk.ncu-rep --page source --csv --print-source sass > gpurun_out/r2_walk_source.csv 2>/dev/null
python tools/ncu_regions.py gpurun_out/r2_walk_source.csv "k_walk_v3<(int)1" > gpurun_out/r2_walk_regions.txt
cp gpurun_out/walk_traffic.json profiles/walk_traffic.json  # bench.py reads it from there (and ignores one from other sources)
python bench.py --steps 4 --warmup 3 2> gpurun_out/r2_bench_n1.err | grep '^{' > gpurun_out/r2_bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/r2_bench_ref.err | grep '^{' > gpurun_out/r2_bench_reference_arm.json
B="--steps 3 --warmup 3 --no-cpu --no-e2e --no-full"
python bench.py --workload c2 $B 2>/dev/null | grep '^{' > gpurun_out/r2_bench_c2.json
python bench.py --workload c2 --repeats 30 $B 2>/dev/null | grep '^{' > gpurun_out/r2_bench_c2_repeats.json
python bench.py --workload c3 $B 2>/dev/null | grep '^{' > gpurun_out/r2_bench_c3_kimura.json
python bench.py --workload c3 --model LOGDET $B 2>/dev/null | grep '^{' > gpurun_out/r2_bench_c3_logdet.json
python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu --no-e2e --no-full 2>/dev/null | grep '^{' > gpurun_out/r2_bench_c5.json
python bench.py --contigs 50 --steps 2 --warmup 3 --no-cpu --no-e2e --no-full 2>/dev/null | grep '^{' > gpurun_out/r2_bench_c4_join50.json
python bench.py --repeats 30 $B 2>/dev/null | grep '^{' > gpurun_out/r2_bench_c4_repeats.json
python bench.py --genomes 512 --repeats 30 $B 2>/dev/null | grep '^{' > gpurun_out/r2_bench_512_repeats.json
python bench.py --genomes 512 $B 2>/dev/null | grep '^{' > gpurun_out/r2_bench_512.json
python bench.py --genomes 512 --divergence 1e-5,1e-4 $B 2>/dev/null | grep '^{' > gpurun_out/r2_bench_512_outbreak.json
python bench.py --genomes 512 --divergence 5e-4,2e-3 $B 2>/dev/null | grep '^{' > gpurun_out/r2_bench_512_lowd.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_bench_*.json")):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f.split("/")[-1], round(d["value"]), round(d["ms_per_step"], 2), d.get("e2e") and round(d["e2e"]["value"]), d.get("roofline") and (d["roofline"]["launch_ms"], d["roofline"]["traffic"], round(d["roofline"]["frac"], 4)), d.get("parity"), d.get("full_matrix") and d["full_matrix"]["seconds"], d.get("esa_build") and round(d["esa_build"]["ms_per_subject"], 3), d.get("cub_calls"))
PY
