"""BASELINE config 4 (scaled): SLQ log-det + gradient on the n = 1M sparse operand, Krylov depth
100, Rademacher probes sharded over the GPUs (torchrun, one rank per GPU), one all-reduce at the end.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_slq.py
Prints one JSON line (rank 0): probes/s, log-det estimate, gradient norm."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import experiments_lanczos_adjoints_b200 as bl
from experiments_lanczos_adjoints_b200 import parallel, synthetic

rank, world, local_rank = parallel.init_from_env()
n = int(os.environ.get("N", 1_000_000))
K = int(os.environ.get("DEPTH", 100))
per_gpu = int(os.environ.get("PROBES_PER_GPU", 8))
dtype = np.float32
row, col, data = synthetic.banded_spd_coo(n, bands=5, seed=0)
num = per_gpu * world
probes = (np.random.default_rng(1).integers(0, 2, size=(num, n), dtype=np.int8) * 2 - 1).astype(dtype)

op = bl.operators.SparseOperator(row, col, (n, n))
integrand = bl.lanczos.integrand_spd(np.log, K, op)
estimate = parallel.hutchinson_sharded(integrand, lambda key: probes)
params = bl.asarray(data.astype(dtype))

value, (grad,) = estimate.value_and_grad(None, params)  # warm-up (allocations, first launches)
bl.synchronize()
t0 = time.perf_counter()
value, (grad,) = estimate.value_and_grad(None, params)
bl.synchronize()
dt = time.perf_counter() - t0
if rank == 0:
    g = grad.numpy()
    print(json.dumps({"world": world, "n": n, "K": K, "probes": num, "seconds": dt, "probes_per_s": num / dt,
                      "krylov_steps_per_s": num * K / dt, "logdet_estimate": float(value),
                      "grad_norm": float(np.linalg.norm(g)), "grad_finite": bool(np.isfinite(g).all())}))
from experiments_lanczos_adjoints_b200 import comm as bl_comm

bl_comm.default().barrier()
bl_comm.shutdown()
