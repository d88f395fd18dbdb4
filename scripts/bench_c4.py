"""BASELINE config 4 through the estimator API with lazily drawn, sharded probes (`bench_extra.slq_probe_sharding`):
python scripts/bench_c4.py [num_probes]   (one GPU; under torchrun: probe sharding)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_extra  # noqa: E402
import experiments_lanczos_adjoints_b200 as bl  # noqa: E402
from experiments_lanczos_adjoints_b200 import comm as bl_comm, synthetic  # noqa: E402

group = bl_comm.init_from_env()
n, K = 1_000_000, 100
row, col, data = synthetic.banded_spd_coo(n, 5, seed=0)
num = int(sys.argv[1]) if len(sys.argv) > 1 else 256
bench_extra.slq_probe_sharding(bl, group, row, col, data, n, K, num_probes=16 * group.world)  # warm-up
out = bench_extra.slq_probe_sharding(bl, group, row, col, data, n, K, num_probes=num)
if group.rank == 0:
    print(json.dumps(out))
bl_comm.shutdown()
