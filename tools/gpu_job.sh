#!/bin/bash
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_walk|k_lcp|k_plcp|k_phi|k_round|k_apply|k_head|k_bucket" -c 60 --csv --log-file gpurun_out/r2_launches_repeats.csv python bench.py --genomes 512 --repeats 30 --steps 1 --warmup 0 --rows 2 --no-cpu --no-e2e --no-full > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r2_launches_repeats.csv")) if len(r)>5]
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
for k,a in sorted(agg.items(), key=lambda x:-x[1][1])[:14]:
    print(f"  {k:60s} {a[0]:4d} {a[1]:10.1f} per-launch {a[1]/a[0]:8.1f}")
PY
