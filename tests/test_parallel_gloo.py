"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: probe partition and the single
all-reduce of the sharded Hutchinson estimator.  The integrand is a stand-in with the same
protocol as `lanczos.integrand_spd` (callable + value_and_grad): the device path needs a GPU
and is covered by the `-m gpu` tests; here only the sharding / reduction plumbing is tested."""

import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest
from conftest import ROOT

from experiments_lanczos_adjoints_b200 import parallel


@pytest.mark.parametrize("num,world", [(1024, 8), (10, 4), (3, 8), (0, 2), (7, 1)])
def test_shard_bounds_partition_exactly(num, world):
    blocks = [parallel.shard_bounds(num, r, world) for r in range(world)]
    assert blocks[0][0] == 0 and blocks[-1][1] == num
    for (lo, hi), (lo2, _) in zip(blocks, blocks[1:]):
        assert hi == lo2 and hi >= lo
    sizes = [hi - lo for lo, hi in blocks]
    assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("n,world", [(10, 2), (11, 4), (5, 8)])
def test_row_shard_index_work_covers_every_entry_once(n, world):
    """Row-sharded sparse operand: every COO entry lands in exactly one rank's A block and one
    rank's A^T block; local indices map back to the global ones (bit-exact host work)."""
    rng = np.random.default_rng(n)
    row, col = rng.integers(0, n, 40), rng.integers(0, n, 40)
    seen_a, seen_b = np.zeros(40, int), np.zeros(40, int)
    for r in range(world):
        chunk, idx_a, row_a, col_a, idx_b, row_b, col_b = parallel.shard_coo_rows(row, col, n, r, world)
        assert chunk * world >= n
        seen_a[idx_a] += 1
        seen_b[idx_b] += 1
        assert np.array_equal(row_a + r * chunk, row[idx_a]) and np.array_equal(col_a, col[idx_a])
        assert np.array_equal(row_b + r * chunk, col[idx_b]) and np.array_equal(col_b, row[idx_b])
        assert row_a.max(initial=0) < chunk and row_b.max(initial=0) < chunk
    assert np.all(seen_a == 1) and np.all(seen_b == 1)


WORKER = textwrap.dedent(
    """
    import os, sys
    import numpy as np
    sys.path.insert(0, os.environ["BL_ROOT"])
    import torch.distributed as dist
    from experiments_lanczos_adjoints_b200 import parallel, hutchinson

    rank, world, _ = parallel.init_from_env(backend="gloo")
    assert world == 2

    class Quadform:  # stand-in integrand: v^T diag(p) v, gradient v*v
        def __call__(self, v, p):
            return float(np.dot(v, p * v))
        def value_and_grad(self, v, p, want_dv0=True):
            return float(np.dot(v, p * v)), (None, v * v)

    rng = np.random.default_rng(0)
    probes = rng.integers(0, 2, size=(7, 5)) * 2.0 - 1.0   # odd count: ragged shards
    p = rng.uniform(1.0, 2.0, 5)
    est = parallel.hutchinson_sharded(Quadform(), lambda key: probes)
    ref = hutchinson.hutchinson(Quadform(), lambda key: probes)
    val = est(None, p)
    val2, (grad,) = est.value_and_grad(None, p)
    rval, (rgrad,) = ref.value_and_grad(None, p)
    assert np.allclose(val, rval) and np.allclose(val2, rval), (val, val2, rval)
    assert np.allclose(grad, rgrad), (grad, rgrad)
    # more ranks than probes: one rank has an empty shard
    one = probes[:1]
    est1 = parallel.hutchinson_sharded(Quadform(), lambda key: one)
    v1, (g1,) = est1.value_and_grad(None, p)
    assert np.allclose(v1, np.dot(one[0], p * one[0])) and np.allclose(g1, one[0] ** 2)
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")
    """
)


def test_sharded_hutchinson_matches_single_process(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, BL_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE="2")
    procs = [
        subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        for r in range(2)
    ]  # fmt: skip
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, out
        assert f"rank {r} ok" in out
