#!/bin/bash
# On the GPU box: one `ncu --set full` capture of the walk kernels of ONE subject of the benched
# workload (3085 x 2.1 Mbp), then the DRAM traffic of that launch group as JSON, tagged with the
# digest of the kernel sources it was taken from (bench.py refuses a file from other sources).
#   tools/capture_traffic.sh <tag>   -> gpurun_out/<tag>_walk.ncu-rep, gpurun_out/walk_traffic.json
tag=${1:-traffic}
ncu --set full --import-source on --clock-control none -k regex:"k_walk_v3|k_walk_reduce" -c 3 -f -o gpurun_out/${tag}_walk \
	python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --no-full --rows 1 > gpurun_out/${tag}_ncu.log 2>&1
ncu -i gpurun_out/${tag}_walk.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
python - "$tag" <<'PY'
import csv, json, sys
sys.path.insert(0, ".")
import bench

tag = sys.argv[1]
rows = list(csv.reader(open(f"gpurun_out/{tag}_raw.csv")))
h, units = rows[0], rows[1]
def val(r, name):
    i = h.index(name)
    v = float(r[i].replace(",", ""))
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(units[i], 1.0)
out = {"capture": f"ncu --set full --clock-control none, bench.py --rows 1 (one subject x 3084 queries), tools/capture_traffic.sh {tag}",
       "kernel_sha": bench.kernel_sources_sha(), "pairs_per_launch": 3084, "kernels": [], "dram_bytes_read": 0.0, "dram_bytes_write": 0.0}
for r in rows[2:]:
    rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
    out["kernels"].append({"name": r[h.index("Kernel Name")].split("(")[0], "dram_bytes_read": rd, "dram_bytes_write": wr,
                           "duration_ms": float(r[h.index("gpu__time_duration.sum")].replace(",", "")) * {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(units[h.index("gpu__time_duration.sum")], 1)})
    out["dram_bytes_read"] += rd
    out["dram_bytes_write"] += wr
json.dump(out, open("gpurun_out/walk_traffic.json", "w"), indent=1)
print(json.dumps(out)[:600])
PY
