#!/bin/bash
# two GPUs: the torch-free multi-GPU layer (socket rendezvous + bl_dist_nccl_*)
O=gpurun_out/r2c7; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_reference_scenarios.py -x -q -k "suitesparse or batched_initial" > $O/tests_new.log 2>&1; echo "exit=$?" >> $O/tests_new.log; tail -15 $O/tests_new.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
BL_BENCH_EXTRA_QUICK=1 timeout 900 $TR --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"; tail -3 $O/bench_n2.err; head -c 1500 $O/bench_n2.json; echo
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2c7/bench_n2.json')); print(json.dumps(d.get('extra'))[:3000])
except Exception as e: print('no json', e)
PY
N=200000 DEPTH=20 timeout 300 $TR --master-port 29512 scripts/run_row_sharded.py > $O/row_sharded_nccl.json 2> $O/row_sharded_nccl.err; echo "row sharded nccl rc=$?"; cat $O/row_sharded_nccl.json; tail -3 $O/row_sharded_nccl.err
GRID=1024 ROUTE=nccl timeout 300 $TR --master-port 29513 scripts/bench_row_sharded_wave.py > $O/wave_nccl.json 2> $O/wave_nccl.err; echo "wave nccl rc=$?"; cat $O/wave_nccl.json; tail -3 $O/wave_nccl.err
GRID=4096 ROUTE=peer timeout 300 $TR --master-port 29514 scripts/bench_row_sharded_wave.py > $O/wave_peer.json 2> $O/wave_peer.err; echo "wave peer rc=$?"; cat $O/wave_peer.json; tail -3 $O/wave_peer.err
