#!/bin/bash
# per-kernel times of build + refit at a given heightfield size (default 2237 -> 10 M triangles), 30- and 63-bit keys
N=${BIG_N:-2237}
for bits in 30 63; do
PROBE_BITS=$bits timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${N}_${bits}.csv python scripts/scale_probe.py $N > /dev/null 2>&1
python - <<PY
import csv, re
lines=[l for l in open('gpurun_out/launches_${N}_${bits}.csv') if not l.startswith('==')]
seen={}
for row in csv.DictReader(lines):
    if row.get('Metric Name')=='gpu__time_duration.sum':
        name=re.sub(r'\(.*','',row['Kernel Name']).replace('void <unnamed>::','').replace('<unnamed>::','').replace('void ','')
        v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
        v = v/1000 if u=='ns' else (v*1000 if u=='ms' else v)
        seen.setdefault(name,[]).append(v)
print("N=${N} bits=${bits}")
for k,v in seen.items():
    print(f"{k:45s} n={len(v):3d}  median {sorted(v)[len(v)//2]:10.1f} us")
PY
PROBE_BITS=$bits python scripts/scale_probe.py $N
done
