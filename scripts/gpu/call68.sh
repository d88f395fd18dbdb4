#!/bin/bash
O=gpurun_out/r2c68; mkdir -p $O
timeout 300 python scripts/profile_reortho_none.py > $O/plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file $O/l.csv python scripts/profile_reortho_none.py > $O/ncu.log 2>&1; echo rc=$?
python - <<'PY'
import csv, collections, re
rows=[r for r in csv.reader(open("gpurun_out/r2c68/l.csv")) if len(r)>5]
h=rows[0]; ki,vi=h.index("Kernel Name"),h.index("Metric Value")
agg=collections.OrderedDict()
for r in rows[1:]:
    n=re.sub(r"\(.*","",r[ki]).replace("void ","")[:56]
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=float(r[vi].replace(",",""))/1e3
tot=sum(a[1] for a in agg.values())
print("launches", sum(a[0] for a in agg.values()), "total ms", round(tot/1e3,2))
for k,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:12]:
    print(f"{k:58s} {c:6d} {t/1e3:8.3f} ms {100*t/tot:5.1f}%  avg {t/c:8.2f} us")
PY
