#!/bin/bash
O=gpurun_out/r2c47; mkdir -p $O
for cfg in "BL_STEP_L2=9" "BL_STEP_L2=13" "BL_STEP_L2=11" "BL_STEP_L2=9 BL_STEP_DEPTH=12" "BL_STEP_L2=13 BL_STEP_DEPTH=12" "BL_STEP_L2=25" "BL_STEP_L2=57"; do
  env $cfg TRACE_PROBES=4 timeout 300 python scripts/trace_step_kernel.py > $O/t.json 2>$O/trace.err
  python - "$cfg" <<'PY'
import json,sys
d=json.load(open("gpurun_out/r2c47/t.json"))
print(sys.argv[1], "S first %.1f s50 %.1f s99 %.1f mean %.1f | P0rest first %.1f mean %.1f | body mean %.1f e2e %.1f"%(d['phases_us_first_step']['phase0 loads'],d['phases_us_step_50']['phase0 loads'],d['phases_us_step_99']['phase0 loads'],d['phases_us_mean']['phase0 loads'],d['phases_us_first_step']['phase0 reduce+epilogue'],d['phases_us_mean']['phase0 reduce+epilogue'],d['kernel_body_us_mean'],d['entry_to_entry_us_mean']))
PY
done
