#!/bin/bash
timeout 600 python scripts/check_grad_tma.py 2>&1 | grep -v Warn | tail -6
for cfg in "BL_GRAD_TMA=0" "BL_GRAD_TMA=1"; do echo "== $cfg tight band"; env $cfg timeout 300 python scripts/time_grad_batch.py f32 tight 2>&1 | grep "P="; env $cfg timeout 300 python scripts/time_grad_batch.py f64 tight 2>&1 | grep "P=";  done
echo "== headline operand (gate: old kernel)"; timeout 300 python scripts/time_grad_batch.py f64 2>&1 | grep "P="
