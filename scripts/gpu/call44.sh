#!/bin/bash
O=gpurun_out/r2c44; mkdir -p $O
for u in 1 2 4; do echo "== BL_GRAD_U=$u"; BL_GRAD_U=$u timeout 300 python scripts/time_grad_batch.py f32 2>&1 | grep "P="; BL_GRAD_U=$u timeout 300 python scripts/time_grad_batch.py f64 2>&1 | grep "P="; done
echo "== tight band, gather kernel"; BL_GRAD_TMA=0 timeout 300 python scripts/time_grad_batch.py f32 tight 2>&1 | grep "P=4"
timeout 600 python scripts/check_grad_tma.py 2>&1 | grep -v Warn | tail -4
for u in 2 4; do BL_GRAD_U=$u timeout 300 python bench.py --quick --steps 6 --warmup 3 > $O/q.json 2>$O/q.err; echo "BL_GRAD_U=$u lockstep: $(cat $O/q.json)"; done
BL_GRAD_U=2 timeout 300 python bench.py --quick --mode streams --probes 1 --steps 5 --warmup 3 > $O/q.json 2>$O/q.err; echo "single: $(cat $O/q.json)"
