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


# ---------------------------------------------------------------------------------------------
# Row sharding: one large operator split by rows over the GPUs (second half)
# ---------------------------------------------------------------------------------------------
class _RawDeviceBuffer:
    """`__cuda_array_interface__` view of a raw device pointer (for torch.as_tensor)."""

    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}  # fmt: skip


def _as_torch(ptr, count, dtype):
    import torch

    typestr = np.dtype(dtype).str
    return torch.as_tensor(_RawDeviceBuffer(ptr, count, typestr), device="cuda")


class PeerComm:
    """Peer-memory communicator (`bl_dist_comm_*`): every rank's mailbox is mapped into every other
    rank's address space (CUDA IPC over NVLink / NVSwitch); reductions and halo exchanges of the
    row-sharded path are then single-block kernels with no NCCL launch and no host callback.
    `torch.distributed` is used once, to exchange the 64-byte IPC handles and for the barrier."""

    def __init__(self, group=None, *, rank=None, world=None):
        import ctypes as C

        from experiments_lanczos_adjoints_b200 import _lib

        self._lib = _lib
        dist = _dist()
        self.group = group
        self.rank = (dist.get_rank(group) if dist else 0) if rank is None else int(rank)
        self.world = (dist.get_world_size(group) if dist else 1) if world is None else int(world)
        h = C.c_void_p()
        _lib.call("bl_dist_comm_create", self.rank, self.world, C.byref(h))
        self.handle = h.value
        self._window_bytes = 0
        self._multi_process = rank is None
        if rank is None and self.world > 1:  # one process per rank: exchange IPC handles
            buf = C.create_string_buffer(64)
            _lib.call("bl_dist_comm_local", self.handle, None, buf)
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(buf.raw), group=group)
            _lib.call("bl_dist_comm_connect_ipc", self.handle, b"".join(handles))
            dist.barrier(group=group)  # every mailbox is zeroed and mapped before the first store

    @property
    def mailbox(self) -> int:
        import ctypes as C

        p = C.c_void_p()
        self._lib.call("bl_dist_comm_local", self.handle, C.byref(p), None)
        return p.value

    def connect_local(self, comms):
        """Same-process ranks (tests: one host thread per rank): connect by mailbox pointer."""
        import ctypes as C

        ptrs = (C.c_void_p * self.world)(*[c.mailbox for c in comms])
        self._lib.call("bl_dist_comm_connect_ptrs", self.handle, ptrs)

    def create_window(self, slot_bytes: int):
        """All-gather window (`bl_dist_comm_window_*`): call on every rank with the same size; with one
        process per rank the IPC handles are exchanged here, same-process ranks call `connect_windows`."""
        import ctypes as C

        if self._window_bytes:
            if slot_bytes > self._window_bytes:
                raise ValueError("the communicator's all-gather window is smaller than requested")
            return
        buf = C.create_string_buffer(64)
        self._lib.call("bl_dist_comm_window_create", self.handle, int(slot_bytes), buf)
        self._window_bytes = int(slot_bytes)
        dist = _dist()
        if self._multi_process and self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(buf.raw), group=self.group)
            self._lib.call("bl_dist_comm_window_connect_ipc", self.handle, b"".join(handles))
            dist.barrier(group=self.group)

    @property
    def window(self) -> int:
        import ctypes as C

        p = C.c_void_p()
        self._lib.call("bl_dist_comm_window_local", self.handle, C.byref(p))
        return p.value

    def connect_windows(self, comms):
        import ctypes as C

        ptrs = (C.c_void_p * self.world)(*[c.window for c in comms])
        self._lib.call("bl_dist_comm_window_connect_ptrs", self.handle, ptrs)

    def timed_out(self) -> bool:
        import ctypes as C

        flag = C.c_int(0)
        self._lib.call("bl_dist_comm_error", self.handle, C.byref(flag))
        return bool(flag.value)

    def close(self):
        h, self.handle = self.handle, None
        if h:
            self._lib.call("bl_dist_comm_destroy", h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class row_sharded:
    """Context manager: while active, every reduction of the Krylov loops on this thread is
    summed over the ranks — the Arnoldi / Lanczos calls then operate on the LOCAL rows of every
    vector and produce the same `H`, coefficients and scalars on every rank.

    `comm=PeerComm(...)`: native route, one single-block peer-memory kernel per reduction
    (`bl_dist_comm_activate`).  Otherwise a hook (`bl_dist_set_reduce_hook`) enqueues an NCCL
    all-reduce on the library's stream.  Neither synchronises the host."""

    def __init__(self, group=None, comm=None):
        from experiments_lanczos_adjoints_b200 import _lib

        self.group = group
        self.comm = comm
        self._lib = _lib

        def hook(_user, values, count, stream):
            try:
                import torch.distributed as dist

                if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
                    return 0  # one rank: the local sums are the global sums
                import torch

                with torch.cuda.stream(torch.cuda.ExternalStream(stream)):
                    dist.all_reduce(_as_torch(values, count, np.float64), group=self.group)
                return 0
            except Exception as exc:  # pragma: no cover - surfaced as BL_ECALLBACK
                self.error = exc
                return 1

        self.error = None
        self._cb = _lib.ALLREDUCE_CB(hook)

    def __enter__(self):
        if self.comm is not None:
            self._lib.call("bl_dist_comm_activate", self.comm.handle)
        else:
            self._lib.call("bl_dist_set_reduce_hook", self._cb, None)
        return self

    def __exit__(self, *exc):
        import ctypes as C

        if self.comm is not None:
            self._lib.call("bl_dist_comm_activate", None)
        else:
            self._lib.call("bl_dist_set_reduce_hook", C.cast(None, self._lib.ALLREDUCE_CB), None)
        return False


class _NativeShardedSparse:
    """Host handle of `bl_op_sharded_sparse_create` (a plain operand of the Krylov calls)."""

    def __new__(cls, handle, chunk, nnz_a, nnz_b):
        from experiments_lanczos_adjoints_b200 import operators as ops

        class Native(ops.Operator):
            def param_shapes(self_inner):
                return [(nnz_a,), (nnz_b,)]

        return Native(handle, chunk)


def shard_coo_rows(row, col, n, rank, world, align=1):
    """Index work of the row-sharded sparse operand (pure host, bit-exact, testable on CPU).

    Uniform chunks of `chunk = ceil(n / world)` rows; rank r owns global rows
    `[r*chunk, min(n, (r+1)*chunk))` and a gathered vector has `world*chunk` entries whose index
    is the global row index.  Returns `(chunk, idx_a, row_a, col_a, idx_b, row_b, col_b)`:
    `idx_a` = COO positions of the entries in the local rows of A (local row index `row_a`, global
    column `col_a`); `idx_b` / `row_b` / `col_b` the same for the local rows of A^T."""
    row = np.asarray(row, dtype=np.int64)
    col = np.asarray(col, dtype=np.int64)
    chunk = -(-int(n) // world)
    chunk = -(-chunk // align) * align  # the peer-memory all-gather moves whole 128-byte lines
    lo, hi = min(int(n), rank * chunk), min(int(n), (rank + 1) * chunk)
    idx_a = np.nonzero((row >= lo) & (row < hi))[0]
    idx_b = np.nonzero((col >= lo) & (col < hi))[0]
    return (chunk, idx_a, (row[idx_a] - lo).astype(np.int32), col[idx_a].astype(np.int32),
            idx_b, (col[idx_b] - lo).astype(np.int32), row[idx_b].astype(np.int32))  # fmt: skip


class RowShardedSparseOperator:
    """The sparse COO operand with its rows sharded over the ranks of `group`.

    `matvec`: all-gather of the local piece of x (NCCL), then the local rows of A;
    `vjp`: all-gather of q and lambda, local rows of A^T for `A^T lambda`, local rows of A for the
    parameter cotangent (`d theta_e` is local to the owner of row_e).  Built on two rectangular
    `SparseOperator`s; plugs into the Krylov loops as a `CallbackOperator`."""

    def __init__(self, row, col, n, group=None, comm=None):
        """`comm=PeerComm(...)`: native route — the all-gathers are single kernels over the communicator's
        peer-memory window (`bl_op_sharded_sparse_create`), `.callback` is a plain operand whose parameters
        are `local_params(data)` and whose cotangent is the local part (`assemble_grad` puts it back in COO
        order).  Without `comm`: NCCL all-gathers from a host callback."""
        from experiments_lanczos_adjoints_b200 import operators as ops

        dist = _dist()
        self.group = group
        self.comm = comm
        self.rank = comm.rank if comm is not None else (dist.get_rank(group) if dist else 0)
        self.world = comm.world if comm is not None else (dist.get_world_size(group) if dist else 1)
        self.n_global, self.nnz = int(n), len(row)
        (self.chunk, self.idx_a, row_a, col_a, self.idx_b, row_b, col_b) = shard_coo_rows(
            row, col, n, self.rank, self.world, align=32 if comm is not None else 1)  # fmt: skip
        width = self.chunk * self.world
        self.A = ops.SparseOperator(row_a, col_a, (self.chunk, width))
        self.B = ops.SparseOperator(row_b, col_b, (self.chunk, width))
        self._gathered = {}
        if comm is not None:
            import ctypes as C

            from experiments_lanczos_adjoints_b200 import _lib

            comm.create_window(width * 8)
            h = C.c_void_p()
            _lib.call("bl_op_sharded_sparse_create", self.A._handle, self.B._handle, comm.handle, self.chunk, width,
                      C.byref(h))  # fmt: skip
            self.callback = _NativeShardedSparse(h.value, self.chunk, len(self.idx_a), len(self.idx_b))
            return
        self.callback = ops.CallbackOperator(self.chunk, self._matvec, self._vjp, num_params=1)
        self.callback.bind = self._bind
        self.callback.grad_zero = self._grad_zero
        self.callback.grad_export = self._grad_export

    def local_params(self, data):
        """The two parameter arrays of the native operand: values of the local rows of A and of A^T."""
        data = np.asarray(data)
        return data[self.idx_a], data[self.idx_b]

    def assemble_grad(self, g_local, dtype=np.float64):
        """Local cotangent (entries of the local rows, operand order) -> global COO order, summed over ranks."""
        full = np.zeros(self.nnz, dtype=np.float64)
        full[self.idx_a] = np.asarray(g_local.numpy() if isinstance(g_local, dev.DeviceArray) else g_local)
        (full,) = all_reduce_sum([full], self.group)
        return full.astype(dtype)

    # vectors ------------------------------------------------------------------------------
    def local_slice(self, x_global):
        """Local piece (zero-padded to `chunk`) of a global host vector."""
        out = np.zeros(self.chunk, dtype=np.asarray(x_global).dtype)
        lo = self.rank * self.chunk
        hi = min(self.n_global, lo + self.chunk)
        out[: hi - lo] = np.asarray(x_global)[lo:hi]
        return out

    def _gather(self, x_loc, slot):
        full = self._gathered.get((slot, x_loc.dtype.str))
        if full is None:
            full = dev.empty((self.chunk * self.world,), x_loc.dtype)
            self._gathered[(slot, x_loc.dtype.str)] = full
        if self.world == 1:
            from experiments_lanczos_adjoints_b200 import _lib

            _lib.call("bl_memcpy_d2d", full.ptr, x_loc.ptr, x_loc.size * x_loc.dtype.itemsize, dev.default_stream().ptr)
            return full
        import torch
        import torch.distributed as dist

        with torch.cuda.stream(torch.cuda.ExternalStream(dev.default_stream().ptr)):
            dist.all_gather_into_tensor(torch.as_tensor(full, device="cuda"), torch.as_tensor(x_loc, device="cuda"),
                                        group=self.group)  # fmt: skip
        return full

    # operator protocol ----------------------------------------------------------------------
    def _bind(self, params, dtype, stream=None):
        (data,) = params
        data = np.asarray(data)
        self.A.bind((data[self.idx_a],), dtype)
        self.B.bind((data[self.idx_b],), dtype)
        self._dtype = np.dtype(dtype)
        return [data]

    def _matvec(self, x_loc):
        return self.A.matvec(self._gather(x_loc, "x"))

    def _vjp(self, q_loc, lam_loc):
        q_full = self._gather(q_loc, "q")
        lam_full = self._gather(lam_loc, "lam")
        z = self.B.matvec(lam_full)
        self.A.vjp(q_full, lam_loc, want_z=False)
        return z, ()

    def _grad_zero(self, dtype, stream=None):
        self.A.grad_zero(dtype)

    def _grad_export(self, dtype, like=None, stream=None):
        """Parameter cotangent in global COO order (each entry is owned by exactly one rank; one
        all-reduce assembles the full vector)."""
        (g_loc,) = self.A.grad_export(dtype)
        full = np.zeros(self.nnz, dtype=np.float64)
        full[self.idx_a] = g_loc.numpy()
        (full,) = all_reduce_sum([full], self.group)
        return [full.astype(dtype)]


class RowShardedWaveOperator:
    """The wave-stencil operand (BASELINE config 5) with the grid rows sharded over the ranks:
    rank r owns rows `[r*gs, (r+1)*gs)` of `u` and of `du` (local state `[u_slab; du_slab]`).
    Each matvec exchanges ONE grid row with each neighbour; the VJP exchanges the boundary rows of
    `q_u` and `lambda_du`.  `d scale` is local to the owner of the row.

    Two routes.  `comm=PeerComm(...)` (native): the slab operator pushes its boundary rows into the
    neighbours' mailboxes itself (`bl_op_wave_set_comm`); `.callback` is then the plain slab operand,
    its parameter is the LOCAL slab of `scale` (`local_scale`) and its cotangent the local slab of
    `d scale` — nothing crosses the host.  Without `comm`: NCCL send/recv driven from a host callback,
    global `scale` in, global `d scale` out (simple, slow)."""

    def __init__(self, grid, stencil, group=None, comm=None):
        from experiments_lanczos_adjoints_b200 import operators as ops

        dist = _dist()
        self.group = group
        self.comm = comm
        self.rank = comm.rank if comm is not None else (dist.get_rank(group) if dist else 0)
        self.world = comm.world if comm is not None else (dist.get_world_size(group) if dist else 1)
        self.g = int(grid)
        if self.g % self.world:
            raise ValueError(f"grid rows ({self.g}) must be divisible by the number of ranks ({self.world})")
        self.gs = self.g // self.world
        self.has_top, self.has_bot = self.rank > 0, self.rank < self.world - 1
        self.local = ops.WaveStencilOperator(self.g, stencil, rows=self.gs, has_top=self.has_top,
                                             has_bottom=self.has_bot)  # fmt: skip
        if comm is not None:
            from experiments_lanczos_adjoints_b200 import _lib

            _lib.call("bl_op_wave_set_comm", self.local._handle, comm.handle)
            self.callback = self.local
            return
        self.callback = ops.CallbackOperator(self.local.n, self._matvec, self._vjp, num_params=1)
        self.callback.bind = self._bind
        self.callback.grad_zero = lambda dtype, stream=None: self.local.grad_zero(dtype)
        self.callback.grad_export = self._grad_export

    def local_slice(self, state):
        """Local `[u_slab; du_slab]` of a global state `(2, g, g)` (or its ravel)."""
        st = np.asarray(state).reshape(2, self.g, self.g)
        lo, hi = self.rank * self.gs, (self.rank + 1) * self.gs
        return np.concatenate([st[0, lo:hi].ravel(), st[1, lo:hi].ravel()])

    def local_scale(self, scale):
        """Local slab (rows of this rank) of the global parameter field `(g, g)`."""
        sc = np.asarray(scale).reshape(self.g, self.g)
        return np.ascontiguousarray(sc[self.rank * self.gs : (self.rank + 1) * self.gs])

    def _exchange(self, field_ptr, dtype, slot_top, slot_bot):
        """Send my first / last row of `field` to the neighbours, receive theirs into the halo slots."""
        if self.world == 1:
            return
        import torch
        import torch.distributed as dist

        g, gs, item = self.g, self.gs, np.dtype(dtype).itemsize
        ops_ = []
        with torch.cuda.stream(torch.cuda.ExternalStream(dev.default_stream().ptr)):
            if self.has_top:
                ops_.append(dist.P2POp(dist.isend, _as_torch(field_ptr, g, dtype), self.rank - 1, self.group))
                ops_.append(dist.P2POp(dist.irecv, _as_torch(self.local.halo_ptr(slot_top), g, dtype),
                                       self.rank - 1, self.group))  # fmt: skip
            if self.has_bot:
                last = field_ptr + (gs - 1) * g * item
                ops_.append(dist.P2POp(dist.isend, _as_torch(last, g, dtype), self.rank + 1, self.group))
                ops_.append(dist.P2POp(dist.irecv, _as_torch(self.local.halo_ptr(slot_bot), g, dtype),
                                       self.rank + 1, self.group))  # fmt: skip
            for req in dist.batch_isend_irecv(ops_):
                req.wait()

    def _bind(self, params, dtype, stream=None):
        from experiments_lanczos_adjoints_b200 import _lib

        (scale,) = params
        sc = np.asarray(scale, dtype=dtype).reshape(self.g, self.g)
        lo, hi = self.rank * self.gs, (self.rank + 1) * self.gs
        self.local.bind((np.ascontiguousarray(sc[lo:hi]),), dtype)
        self._dtype = np.dtype(dtype)
        # the neighbours' boundary rows of `scale` (the parameter is replicated on the hosts)
        s = dev.default_stream()
        for present, row, slot in ((self.has_top, lo - 1, 4), (self.has_bot, hi, 5)):
            if present:
                host = np.ascontiguousarray(sc[row])
                _lib.call("bl_memcpy_h2d", self.local.halo_ptr(slot), host.ctypes.data, host.nbytes, s.ptr)
                s.synchronize()
        return [scale]

    def _matvec(self, x_loc):
        self._exchange(x_loc.ptr, x_loc.dtype, 0, 1)  # rows of u
        return self.local.matvec(x_loc)

    def _vjp(self, q_loc, lam_loc):
        half = self.gs * self.g * q_loc.dtype.itemsize
        self._exchange(q_loc.ptr, q_loc.dtype, 0, 1)             # rows of q_u
        self._exchange(lam_loc.ptr + half, lam_loc.dtype, 2, 3)  # rows of lambda_du
        return self.local.vjp(q_loc, lam_loc), ()

    def _grad_export(self, dtype, like=None, stream=None):
        (g_loc,) = self.local.grad_export(dtype)
        full = np.zeros((self.g, self.g), dtype=np.float64)
        full[self.rank * self.gs : (self.rank + 1) * self.gs] = g_loc.numpy()
        (full,) = all_reduce_sum([full], self.group)
        return [full.astype(dtype)]
