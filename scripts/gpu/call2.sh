#!/bin/bash
# first run of k_step_tma on a B200: small parity cases under a timeout, then A/B timings
O=gpurun_out/r2c2; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "sparse_tridiag or symmetric_adjoint or golden or nonsymmetric" > $O/tests_step.log 2>&1; echo "exit=$?" >> $O/tests_step.log
tail -3 $O/tests_step.log
if grep -q "exit=0" $O/tests_step.log; then
  for cfg in "BL_STEP=1" "BL_STEP=0" "BL_STEP=1 BL_STEP_PDL=0"; do
    for P in 1 4; do
      tag=$(echo "$cfg" | tr ' =' '__')_p$P
      env $cfg timeout 300 python bench.py --quick --probes $P --steps 10 --warmup 3 > $O/bench_$tag.json 2> $O/bench_$tag.err; echo "$tag rc=$? $(cat $O/bench_$tag.json)"
    done
  done
  env BL_STEP=1 BL_BLOCKS_PER_SM=2 timeout 300 python bench.py --quick --probes 4 --steps 10 --warmup 3 > $O/bench_step_p4_bps2.json 2>/dev/null; echo "p4 bps2 $(cat $O/bench_step_p4_bps2.json)"
  timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > $O/tests_all.log 2>&1; echo "exit=$?" >> $O/tests_all.log
  tail -4 $O/tests_all.log
fi
