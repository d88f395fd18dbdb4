#!/bin/bash
# after the fix (phase S's mbarriers in memory of their own): lanes in flight vs rep 0, gathers through L2 / L1, quick bench lines,
# then the whole GPU suite and the bench line
O=gpurun_out/r2c30; mkdir -p $O
run() { echo "== $1"; shift; env "$@" DBG_REPS=6 timeout 200 python scripts/debug_lockstep_op.py child /tmp/x.npz 1000000 30 4 ${LANES:-2} f32 2>&1 | grep -v Warn | tail -3; }
run "fused 2 lanes, gathers ld.global.cg" BL_STEP_L2=1
run "fused 2 lanes, gathers ld.global.nc" BL_STEP_L2=65
LANES=3 run "fused 3 lanes" BL_STEP_L2=1
for cfg in "BL_STEP_L2=1" "BL_STEP_L2=65" "BL_STEP_OP=0"; do
  env $cfg timeout 300 python bench.py --quick --steps 5 --warmup 3 > $O/q.json 2>$O/q.err; echo "$cfg lockstep: $(cat $O/q.json)"; tail -2 $O/q.err
done
timeout 1500 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "exit=$?" >> $O/tests.log; tail -3 $O/tests.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke exit=$?"; tail -2 $O/smoke.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench exit=$?"; head -c 1500 $O/bench.json; tail -3 $O/bench.err
