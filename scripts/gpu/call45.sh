#!/bin/bash
# after the batched loads in k_sell_grad_batch: GPU suite, bench line, full capture of the kernel (400 pairs at C2)
O=gpurun_out/r2c45; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "exit=$?" >> $O/tests.log; tail -3 $O/tests.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench exit=$?"; tail -3 $O/bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c45/bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "single", d["config"]["single_probe"], "frac", d["roofline"]["frac"], "whole", d["roofline"]["whole_step"]["frac"])
print(json.dumps(d["roofline"]["classes"]))
print(json.dumps(d["extra"]["c4_slq_probe_sharding"]))
print(json.dumps(d["extra"]["published_recipe_c2_dense_cotangent"]))
PY
CMD="python bench.py --quick --steps 1 --warmup 1 --lanes 1 --probes 4"
timeout 300 $CMD > $O/plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sell_grad_batch -s 1 -c 1 -o $O/prof_grad_batch $CMD > $O/ncu_gb.log 2>&1; echo "ncu grad batch rc=$?"
