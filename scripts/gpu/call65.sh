#!/bin/bash
O=gpurun_out/r2c65; mkdir -p $O
timeout 300 python scripts/bench_gp_logml.py > $O/plain.log 2>&1; tail -2 $O/plain.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file $O/l.csv python scripts/bench_gp_logml.py > $O/ncu.log 2>&1; echo rc=$?
python - <<'PY'
import csv, collections, re
rows=[r for r in csv.reader(open("gpurun_out/r2c65/l.csv")) if len(r)>5]
h=rows[0]; ki,vi=h.index("Kernel Name"),h.index("Metric Value")
agg=collections.OrderedDict()
for r in rows[1:]:
    n=re.sub(r"\(.*","",r[ki]).replace("void ","")[:56]
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=float(r[vi].replace(",",""))/1e3
tot=sum(a[1] for a in agg.values())
print("launches", sum(a[0] for a in agg.values()), "total ms", round(tot/1e3,1))
for k,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:16]:
    print(f"{k:58s} {c:6d} {t/1e3:8.2f} ms {100*t/tot:5.1f}%  avg {t/c:8.1f} us")
PY
