#!/bin/bash
O=gpurun_out/r2c46; mkdir -p $O
TRACE_PROBES=4 BL_STEP_L2=9 timeout 300 python scripts/trace_step_kernel.py > $O/trace4.json 2>$O/trace.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2c46/trace4.json"))
for k in ['phases_us_first_step','phases_us_step_50','phases_us_step_99','phases_us_mean']:
    print(k,{a:round(b,1) for a,b in d[k].items()})
print('body mean',round(d['kernel_body_us_mean'],1),'entry-to-entry',round(d['entry_to_entry_us_mean'],1))
PY
for i in 1 2; do timeout 300 python bench.py --quick --steps 6 --warmup 3 > $O/q.json 2>$O/q.err; echo "lockstep: $(cat $O/q.json)"; done
timeout 900 python -m pytest tests -m gpu -x -q -k "lockstep or lanes or slq or headline or golden or nonsymmetric" 2>&1 | tail -2
