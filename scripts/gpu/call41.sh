#!/bin/bash
# 8 GPUs: the bench line as the driver launches it (probe sharding + extra configurations)
O=gpurun_out/r2c41; mkdir -p $O
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err; echo "bench n8 exit=$?"; head -c 400 $O/bench_n8.json; echo; grep -i "warn\|error\|symmetric" $O/bench_n8.err | head -5
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c41/bench_n8.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
for k in ("c4_slq_probe_sharding", "c5_wave_row_sharded"):
    print(k, json.dumps(d["extra"][k]))
PY
