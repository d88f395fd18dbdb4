#!/bin/bash
# 2 GPUs: the bench line as the driver launches it (probe sharding + extra configurations), reference arm skipped
O=gpurun_out/r2c24; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 exit=$?"; cat $O/bench_n2.json | head -c 6000; tail -5 $O/bench_n2.err
