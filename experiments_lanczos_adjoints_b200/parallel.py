"""Multi-GPU layer: one process per GPU, `torch.distributed` for the plumbing.

Probe sharding (BASELINE config 4): the Hutchinson / SLQ probe vectors are
independent Lanczos runs, so each rank takes a contiguous block of the probes, runs forward +
adjoint on its own GPU with a replicated operator, and the estimator ends with ONE all-reduce
of `(sum of quadratic forms, probe count, parameter cotangents)` — NCCL over NVLink on GPUs,
gloo in the CPU tests.  No collective sits on the data path of a probe.

torch is imported lazily and only here: the single-GPU product path does not depend on it.
"""

from __future__ import annotations

import os

import numpy as np

from experiments_lanczos_adjoints_b200 import device as dev
from experiments_lanczos_adjoints_b200.hutchinson import _scale, probe_sum


def shard_bounds(num: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block `[lo, hi)` of `num` units for `rank`; blocks differ by at most one unit
    and cover `range(num)` exactly (ranks beyond `num` get an empty block)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world: {rank}/{world}")
    base, extra = divmod(int(num), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def init_from_env(backend: str | None = None):
    """Initialise `torch.distributed` from the torchrun environment (RANK / WORLD_SIZE /
    LOCAL_RANK / MASTER_ADDR / MASTER_PORT) and bind this process to its GPU.
    Returns `(rank, world, local_rank)`; a single-process run needs no initialisation."""
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        import torch
        import torch.distributed as dist

        if not dist.is_initialized():
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            kwargs = {}
            if backend == "nccl":
                torch.cuda.set_device(local_rank)
                kwargs["device_id"] = torch.device("cuda", local_rank)
            dist.init_process_group(backend, **kwargs)
    if dev.device_count() > 0:
        dev.set_device(local_rank)
    return rank, world, local_rank


def _dist():
    import torch.distributed as dist

    return dist if dist.is_available() and dist.is_initialized() else None


def all_reduce_sum(values, group=None):
    """Sum a list of arrays over the ranks with ONE collective call.

    Host arrays / scalars are packed into one buffer; `DeviceArray`s are reduced in place
    through a zero-copy torch view (`__cuda_array_interface__`).  Returns the reduced list."""
    dist = _dist()
    if dist is None or dist.get_world_size(group) == 1:
        return list(values)
    import torch

    backend = dist.get_backend(group)
    host_idx = [i for i, v in enumerate(values) if not isinstance(v, dev.DeviceArray)]
    out = list(values)
    if host_idx:
        flat = np.concatenate([np.asarray(values[i], dtype=np.float64).reshape(-1) for i in host_idx])
        t = torch.from_numpy(flat.copy())
        if backend == "nccl":
            t = t.cuda()
        dist.all_reduce(t, group=group)
        flat = t.cpu().numpy()
        pos = 0
        for i in host_idx:
            shape = np.shape(values[i])
            size = int(np.prod(shape, dtype=np.int64)) if shape else 1
            out[i] = flat[pos : pos + size].reshape(shape)
            pos += size
    for i, v in enumerate(values):
        if isinstance(v, dev.DeviceArray):
            dev.default_stream().synchronize()  # our kernels run on our own stream
            t = torch.as_tensor(v, device="cuda")
            dist.all_reduce(t, group=group)
            torch.cuda.current_stream().synchronize()
            out[i] = v
    return out


class _ShardedEstimator:
    """`hutchinson.hutchinson(integrand, sample_fun)` with the probes sharded over the ranks."""

    def __init__(self, integrand_fun, sample_fun, group=None):
        self.integrand_fun, self.sample_fun, self.group = integrand_fun, sample_fun, group

    def _local(self, key):
        samples = self.sample_fun(key)  # same key on every rank -> same probe matrix
        dist = _dist()
        rank = dist.get_rank(self.group) if dist else 0
        world = dist.get_world_size(self.group) if dist else 1
        num = samples._shape[0] if isinstance(samples, dev.DeviceArray) else len(samples)
        lo, hi = shard_bounds(num, rank, world)
        if isinstance(samples, dev.DeviceArray):
            return [samples.row(i) for i in range(lo, hi)]
        return np.asarray(samples)[lo:hi]

    def __call__(self, key, *parameters):
        local = self._local(key)
        total, _, count = probe_sum(self.integrand_fun, local, parameters) if len(local) else (0.0, None, 0)
        total, count = all_reduce_sum([np.asarray(total, np.float64), np.asarray(float(count))], self.group)
        return total / count

    def value_and_grad(self, key, *parameters):
        local = self._local(key)
        if len(local):
            total, grads, count = probe_sum(self.integrand_fun, local, parameters, with_grad=True)
        else:  # a rank without probes still takes part in the collective
            total, count = 0.0, 0
            grads = [np.zeros(np.shape(p)) for p in parameters]
        red = all_reduce_sum([np.asarray(total, np.float64), np.asarray(float(count)), *grads], self.group)
        total, count, grads = red[0], red[1], red[2:]
        return total / count, tuple(_scale(g, 1.0 / float(count)) for g in grads)


def hutchinson_sharded(integrand_fun, /, sample_fun, group=None):
    """Probe-sharded Hutchinson estimator: same call signature and (up to summation order) the
    same value / gradient as `hutchinson.hutchinson` on one GPU."""
    return _ShardedEstimator(integrand_fun, sample_fun, group)
