#!/bin/bash
O=gpurun_out/r2c62; mkdir -p $O
timeout 300 python scripts/profile_dense_cotangent.py 2>&1 | grep -v Warn | tail -3
timeout 300 python scripts/profile_dense_cotangent.py > /dev/null 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/l.csv python scripts/profile_dense_cotangent.py > $O/ncu.log 2>&1
python - <<'PY'
import csv, collections, re
rows=[r for r in csv.reader(open("gpurun_out/r2c62/l.csv")) if len(r)>5]
h=rows[0]; ki,vi=h.index("Kernel Name"),h.index("Metric Value")
agg=collections.OrderedDict()
for r in rows[1:]:
    n=re.sub(r"\(.*","",r[ki]).replace("void ","")[:48]
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=float(r[vi].replace(",",""))/1e3
tot=sum(a[1] for a in agg.values())
for k,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:12]:
    print(f"{k:50s} {c:5d} {t/1e3:8.2f} ms {100*t/tot:5.1f}%")
PY
