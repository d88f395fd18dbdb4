#!/bin/bash
# whole GPU suite, smoke, bench line (with extras), reference arm
O=gpurun_out/r2c40; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "exit=$?" >> $O/tests.log; tail -3 $O/tests.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke exit=$?"; tail -2 $O/smoke.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench exit=$?"; head -c 600 $O/bench.json; echo; tail -3 $O/bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c40/bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "single", d["config"]["single_probe"], "frac", d["roofline"]["frac"], "whole", d["roofline"]["whole_step"]["frac"])
print(json.dumps(d["extra"]["c4_slq_probe_sharding"]))
print(json.dumps(d["extra"]["c3_gp"]))
PY
