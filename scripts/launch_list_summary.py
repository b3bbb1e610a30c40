import csv,collections,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
d=collections.defaultdict(list)
for r in rows[1:]:
    try: d[r[ki].replace("void ","").replace("<unnamed>::","")[:50]].append(float(r[vi].replace(",","")))
    except: pass
for k,v in d.items(): print(f"{k:50s} n={len(v):3d} med={sorted(v)[len(v)//2]/1000:9.1f} us")
