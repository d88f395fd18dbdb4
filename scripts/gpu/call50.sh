#!/bin/bash
O=gpurun_out/r2c50; mkdir -p $O
timeout 600 python scripts/check_step_xi.py 2>&1 | grep -v Warn | tail -10
for cfg in "BL_STEP_XI=0" "BL_STEP_XI=1" "BL_STEP_XI=0" "BL_STEP_XI=1"; do
  env $cfg timeout 300 python bench.py --quick --steps 6 --warmup 3 > $O/q.json 2>$O/q.err; echo "$cfg lockstep: $(cat $O/q.json)"; tail -1 $O/q.err
done
for cfg in "BL_STEP_XI=0" "BL_STEP_XI=1"; do
  env $cfg BL_STEP_L2=9 TRACE_PROBES=4 timeout 300 python scripts/trace_step_kernel.py > $O/t.json 2>$O/trace.err
  python - "$cfg" <<'PY'
import json,sys
d=json.load(open("gpurun_out/r2c50/t.json"))
print(sys.argv[1], "S first %.1f s50 %.1f s99 %.1f mean %.1f | P0rest mean %.1f | body mean %.1f e2e %.1f"%(d['phases_us_first_step']['phase0 loads'],d['phases_us_step_50']['phase0 loads'],d['phases_us_step_99']['phase0 loads'],d['phases_us_mean']['phase0 loads'],d['phases_us_mean']['phase0 reduce+epilogue'],d['kernel_body_us_mean'],d['entry_to_entry_us_mean']))
PY
done
