#!/bin/bash
# 2 GPUs: the bench line as the driver launches it (probe sharding + extra configurations) after the step-kernel fix
O=gpurun_out/r2c34; mkdir -p $O
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 exit=$?"; head -c 1200 $O/bench_n2.json; echo; grep -i "warn\|error\|symmetric" $O/bench_n2.err | head -5
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c34/bench_n2.json"))
print(json.dumps(d.get("extra", {}), indent=1)[:6000])
PY
