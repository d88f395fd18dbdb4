#!/bin/bash
for cfg in "BL_GRAD_TMA=0" "BL_GRAD_TMA=1"; do echo "== $cfg"; env $cfg timeout 300 python scripts/time_grad_batch.py f32 2>&1 | grep "P="; done
BL_GRAD_TMA=1 timeout 300 python scripts/time_grad_batch.py f64 2>&1 | grep "P="
BL_GRAD_TMA=0 timeout 300 python scripts/time_grad_batch.py f64 2>&1 | grep "P="
