#!/bin/bash
O=gpurun_out/r2c67; mkdir -p $O
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke exit=$?"; tail -1 $O/smoke.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench exit=$?"; tail -2 $O/bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c67/bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "single", d["config"]["single_probe"]["krylov_steps_per_s"], "frac", d["roofline"]["frac"], "whole", d["roofline"]["whole_step"]["frac"], "stagger", d["config"]["stagger"])
for k, v in d["extra"].items():
    print(k, json.dumps(v)[:420])
PY
