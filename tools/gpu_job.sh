#!/bin/bash
# extended boundary replay + cooperative scan: parity, then the repeats workloads against the plain ones
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py -m gpu -x -q 2>&1 | tail -3
for wl in "--genomes 512 --repeats 30" "--genomes 512" "" "--repeats 30" "--workload c2" "--workload c2 --repeats 30"; do
python bench.py $wl --steps 3 --warmup 2 --no-cpu --no-e2e --no-full 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); n=d['config']['rows_per_step_per_gpu']*d['steps']
print('$wl',round(d['value']),'ms/step',round(d['ms_per_step'],2),'walk',round(d['kernel_ms_sums_of_timed_steps']['walk']/n,3),'esa',round(d['kernel_ms_sums_of_timed_steps']['esa']/n,3),'alone',d['roofline']['launch_ms'], 'esa alone', d['esa_build']['ms_per_subject'], 'rounds', d['esa_build']['sa_rounds_per_subject'])"
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_walk|k_lcp|k_plcp|k_phi|k_round|k_apply|k_head|k_bucket" -c 60 --csv --log-file gpurun_out/r2q_launches_repeats.csv python bench.py --genomes 512 --repeats 30 --steps 1 --warmup 0 --rows 2 --no-cpu --no-e2e --no-full > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r2q_launches_repeats.csv")) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if r[0]=='ID'][0]
h=rows[hdr]; ix={n:i for i,n in enumerate(h)}
agg=collections.OrderedDict()
for r in rows[hdr+2:]:
    try:
        k=r[ix['Kernel Name']].split('(')[0][:60]; v=float(r[ix['Metric Value']].replace(',',''))
    except Exception: continue
    u=r[ix['Metric Unit']]
    v = v/1000 if u=='ns' else v*1000 if u=='ms' else v
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
for k,a in sorted(agg.items(), key=lambda x:-x[1][1])[:12]:
    print(f"  {k:60s} {a[0]:4d} {a[1]:10.1f} per-launch {a[1]/a[0]:8.1f}")
PY
