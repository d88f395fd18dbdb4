#!/bin/bash
O=gpurun_out/r2c57; mkdir -p $O
timeout 300 python scripts/profile_wave.py > $O/plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_wave -c 40 --csv --log-file $O/wave.csv python scripts/profile_wave.py > $O/ncu.log 2>&1; echo rc=$?
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r2c57/wave.csv")) if len(r)>5]
h=rows[0]; ki,mi,vi=h.index("Kernel Name"),h.index("Metric Name"),h.index("Metric Value")
agg=collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows[1:]:
    agg[r[ki][:40]][r[mi]].append(float(r[vi].replace(",","")))
for k,m in agg.items():
    print(k, {a:(round(sum(b)/len(b),1), h[h.index("Metric Unit")] if False else '') for a,b in m.items()}, len(next(iter(m.values()))))
PY
grep -m3 "Metric Unit\|k_wave_vjp_a" gpurun_out/r2c57/wave.csv | head -4
