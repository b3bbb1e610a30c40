#!/bin/bash
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1]) if __import__('os').path.exists('gpurun_out/bench_quick.json') else {}
print('signed', d.get('extra', {}).get('signed'))
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_big.csv python scripts/scale_probe.py ${BIG_N:-7072} > /dev/null 2>&1
python - <<'PY'
import csv, re
lines=[l for l in open('gpurun_out/launches_big.csv') if not l.startswith('==')]
seen={}
for row in csv.DictReader(lines):
    if row.get('Metric Name')=='gpu__time_duration.sum':
        name=re.sub(r'\(.*','',row['Kernel Name']).replace('void <unnamed>::','').replace('<unnamed>::','').replace('void ','')
        v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
        v = v/1000 if u=='ns' else (v*1000 if u=='ms' else v)
        seen.setdefault(name,[]).append(v)
for k,v in seen.items():
    print(f"{k:45s} n={len(v):3d}  median {sorted(v)[len(v)//2]:10.1f} us")
PY
