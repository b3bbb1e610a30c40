#!/bin/bash
mkdir -p gpurun_out
export PROF_NQ=$((1<<20))
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tmp.csv python scripts/prof_driver.py > /dev/null 2>&1
python - <<'PY'
import csv, re
lines=[l for l in open('gpurun_out/launches_tmp.csv') if not l.startswith('==')]
seen={}
for row in csv.DictReader(lines):
    if row.get('Metric Name')=='gpu__time_duration.sum':
        name=re.sub(r'\(.*','',row['Kernel Name']).replace('void <unnamed>::','').replace('<unnamed>::','')
        v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
        v = v/1000 if u=='ns' else (v*1000 if u=='ms' else v)
        seen.setdefault(name,[]).append(v)
for k,v in seen.items():
    print(f"{k:45s} n={len(v):3d}  median {sorted(v)[len(v)//2]:9.1f} us")
PY
