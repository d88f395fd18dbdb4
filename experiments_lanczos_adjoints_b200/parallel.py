"""Multi-GPU layer: one process per GPU, NCCL from the library itself (no PyTorch).

Probe sharding (BASELINE config 4): the Hutchinson / SLQ probe vectors are
independent Lanczos runs, so each rank takes a contiguous block of the probes, runs forward +
adjoint on its own GPU with a replicated operator, and the estimator ends with ONE all-reduce
of `(sum of quadratic forms, probe count, parameter cotangents)` -- `ncclAllReduce` over NVLink
through `bl_dist_nccl_allreduce` when the cotangents live on a GPU, the host communicator's sum in
the CPU tests.  No collective sits on the data path of a probe.

`group` arguments are host communicators (`comm.Comm`: `comm.init_from_env()` builds one from the
torchrun environment; `None` means the process-wide default); `comm` arguments are peer-memory
communicators (`PeerComm`).
"""

from __future__ import annotations

import numpy as np

from experiments_lanczos_adjoints_b200 import comm as _comm
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


def init_from_env():
    """Build the process-wide host communicator from the torchrun environment (RANK / WORLD_SIZE /
    LOCAL_RANK / MASTER_ADDR / MASTER_PORT) and bind this process to its GPU.
    Returns `(rank, world, local_rank)`; a single-process run needs no initialisation."""
    c = _comm.init_from_env()
    return c.rank, c.world, c.local_rank


def _group(group):
    return group if group is not None else _comm.default()


def all_reduce_sum(values, group=None):
    """Sum a list of arrays over the ranks: ONE collective for everything that lives on the host (scalars and
    host arrays are packed into one buffer) and ONE NCCL all-reduce for the device arrays (packed into one device
    buffer when there are several).  Every rank must pass the same kinds and shapes in the same order.  Returns
    the reduced list (device arrays are reduced in place)."""
    g = _group(group)
    out = list(values)
    if g.world == 1:
        return out
    host_idx = [i for i, v in enumerate(values) if not isinstance(v, dev.DeviceArray)]
    dev_idx = [i for i, v in enumerate(values) if isinstance(v, dev.DeviceArray)]
    if host_idx:
        flat = np.concatenate([np.asarray(values[i], dtype=np.float64).reshape(-1) for i in host_idx])
        flat = g.allreduce_host(flat)
        pos = 0
        for i in host_idx:
            shape = np.shape(values[i])
            size = int(np.prod(shape, dtype=np.int64)) if shape else 1
            out[i] = flat[pos : pos + size].reshape(shape)
            pos += size
    if dev_idx:
        stream = dev.default_stream()
        dtypes = {values[i].dtype for i in dev_idx}
        contiguous = all(values[i].ld in (None, values[i]._shape[-1]) or values[i].ndim == 1 for i in dev_idx)
        if len(dev_idx) == 1 or len(dtypes) > 1 or not contiguous:
            for i in dev_idx:  # one array (the usual case: one parameter) or mixed dtypes: reduce in place
                g.allreduce_device(values[i].ptr, values[i].size, values[i].dtype, stream)
        else:  # several parameters: pack, one all-reduce, unpack
            from experiments_lanczos_adjoints_b200 import _lib

            dtype = dtypes.pop()
            item = dtype.itemsize
            pack = dev.DeviceArray((sum(values[i].size for i in dev_idx),), dtype)
            pos = 0
            for i in dev_idx:
                _lib.call("bl_memcpy_d2d", pack.ptr + pos * item, values[i].ptr, values[i].size * item, stream.ptr)
                pos += values[i].size
            g.allreduce_device(pack.ptr, pack.size, dtype, stream)
            pos = 0
            for i in dev_idx:
                _lib.call("bl_memcpy_d2d", values[i].ptr, pack.ptr + pos * item, values[i].size * item, stream.ptr)
                pos += values[i].size
        stream.synchronize()
    return out


class _ShardedEstimator:
    """`hutchinson.hutchinson(integrand, sample_fun)` with the probes sharded over the ranks."""

    def __init__(self, integrand_fun, sample_fun, group=None):
        self.integrand_fun, self.sample_fun, self.group = integrand_fun, sample_fun, group

    def _local(self, key):
        """This rank's block of the probes.  A sampler that can generate a slice (`sample_fun(key, lo, hi)`,
        e.g. `parallel.sharded_sampler`) is asked for the local block only; a plain `sample_fun(key)` is called
        as the reference calls it (same key on every rank -> same probe matrix) and sliced."""
        g = _group(self.group)
        slicer = getattr(self.sample_fun, "sample_slice", None)
        if slicer is not None:
            lo, hi = shard_bounds(self.sample_fun.num, g.rank, g.world)
            return slicer(key, lo, hi)
        samples = self.sample_fun(key)
        num = samples._shape[0] if isinstance(samples, dev.DeviceArray) else len(samples)
        lo, hi = shard_bounds(num, g.rank, g.world)
        if isinstance(samples, dev.DeviceArray):
            return [samples.row(i) for i in range(lo, hi)]
        return np.asarray(samples)[lo:hi]

    def __call__(self, key, *parameters):
        local = self._local(key)
        total, _, count = probe_sum(self.integrand_fun, local, parameters) if len(local) else (0.0, None, 0)
        total, count = all_reduce_sum([np.asarray(total, np.float64), np.asarray(float(count))], self.group)
        return total / count

    def _zero_grads(self, parameters):
        """What a rank without probes contributes: zeros of the kind and shape the other ranks' gradients have
        (device arrays in the operator's exported shapes when the integrand runs on the device), so that every
        rank issues the same collectives with the same element counts."""
        op = getattr(getattr(getattr(self.integrand_fun, "alg", None), "alg", None), "op", None)
        if op is not None and hasattr(op, "param_shapes") and dev.device_count() > 0:
            dtype = next((np.asarray(p).dtype for p in parameters if np.asarray(p).dtype.kind == "f"), np.dtype(np.float64))
            if isinstance(parameters[0], dev.DeviceArray):
                dtype = parameters[0].dtype
            return [dev.zeros(tuple(s) or (1,), dtype) for s in op.param_shapes()]
        return [np.zeros(np.shape(p)) for p in parameters]

    def value_and_grad(self, key, *parameters):
        local = self._local(key)
        if len(local):
            total, grads, count = probe_sum(self.integrand_fun, local, parameters, with_grad=True)
        else:  # a rank without probes still takes part in the collectives, with matching buffers
            total, count = 0.0, 0
            grads = self._zero_grads(parameters)
        red = all_reduce_sum([np.asarray(total, np.float64), np.asarray(float(count)), *grads], self.group)
        total, count, grads = red[0], red[1], red[2:]
        return total / count, tuple(_scale(g, 1.0 / float(count)) for g in grads)


def hutchinson_sharded(integrand_fun, /, sample_fun, group=None):
    """Probe-sharded Hutchinson estimator: same call signature and (up to summation order) the
    same value / gradient as `hutchinson.hutchinson` on one GPU."""
    return _ShardedEstimator(integrand_fun, sample_fun, group)


class LazyProbes:
    """`(num, n)` probe matrix whose rows are drawn when they are sliced: the estimator takes four rows at a time,
    so the host generates the next batch while the GPU runs the previous one, and no rank ever holds more than a
    batch (1024 x 1M float32 probes would be 4 GB)."""

    ndim = 2

    def __init__(self, make_rows, lo, hi, n, dtype):
        self._make, self._lo, self._hi = make_rows, int(lo), int(hi)
        self.shape, self.dtype = (max(0, self._hi - self._lo), int(n)), np.dtype(dtype)

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            start, stop, step = idx.indices(len(self))
            if step != 1:
                raise IndexError("LazyProbes supports contiguous slices")
            return self._make(self._lo + start, self._lo + max(start, stop))
        if idx < 0:
            idx += len(self)
        if not 0 <= idx < len(self):
            raise IndexError(idx)
        return self._make(self._lo + idx, self._lo + idx + 1)[0]

    def __array__(self, dtype=None, copy=None):
        out = self._make(self._lo, self._hi)
        return out if dtype is None else out.astype(dtype)


class sharded_sampler:
    """Rademacher probes that every rank can generate BY SLICE: probe i is drawn from its own child stream of the
    key, so rank r materialises only its block `[lo, hi)` (1024 x 1M fp32 probes are 4 GB: not on every rank)
    and the union over ranks is the same `(num, n)` matrix whatever the number of ranks."""

    def __init__(self, x_like, /, *, num: int):
        self.n, self.dtype, self.num = int(np.size(x_like)), np.asarray(x_like).dtype, int(num)

    def _rows(self, key, lo, hi):
        # child stream i of the key, as `hutchinson.split(key, num)[i]` derives it (only the children that are needed:
        # the estimator asks for four rows at a time); one random BIT per entry -- the host draws a batch in a few
        # milliseconds, under the other lane's kernels
        if not isinstance(key, np.random.SeedSequence):
            key = np.random.SeedSequence(int(np.asarray(key).sum()))
        out = np.empty((max(0, hi - lo), self.n), dtype=self.dtype)
        nbytes = (self.n + 7) // 8
        for i in range(lo, hi):
            child = np.random.SeedSequence(entropy=key.entropy, spawn_key=tuple(key.spawn_key) + (i,))
            bits = np.unpackbits(np.frombuffer(np.random.default_rng(child).bytes(nbytes), np.uint8))[: self.n]
            row = out[i - lo]
            row[:] = bits
            row *= 2
            row -= 1
        return out

    def sample_slice(self, key, lo, hi):
        """Rows `[lo, hi)` of the probe matrix, drawn lazily (`LazyProbes`)."""
        return LazyProbes(lambda a, b: self._rows(key, a, b), lo, hi, self.n, self.dtype)

    def __call__(self, key):
        return self._rows(key, 0, self.num)


# ---------------------------------------------------------------------------------------------
# Row sharding: one large operator split by rows over the GPUs (second half)
# ---------------------------------------------------------------------------------------------
class PeerComm:
    """Peer-memory communicator (`bl_dist_comm_*`): every rank's mailbox is mapped into every other
    rank's address space (CUDA IPC over NVLink / NVSwitch); reductions and halo exchanges of the
    row-sharded path are then single-block kernels with no NCCL launch and no host callback.
    The host communicator (`group`) is used once, to exchange the 64-byte IPC handles and for the barrier."""

    def __init__(self, group=None, *, rank=None, world=None):
        import ctypes as C

        from experiments_lanczos_adjoints_b200 import _lib

        self._lib = _lib
        g = _group(group)
        self.group = group
        self.rank = g.rank if rank is None else int(rank)
        self.world = g.world if world is None else int(world)
        h = C.c_void_p()
        _lib.call("bl_dist_comm_create", self.rank, self.world, C.byref(h))
        self.handle = h.value
        self._window_bytes = 0
        self._multi_process = rank is None
        if rank is None and self.world > 1:  # one process per rank: exchange IPC handles
            buf = C.create_string_buffer(64)
            _lib.call("bl_dist_comm_local", self.handle, None, buf)
            handles = g.allgather_bytes(bytes(buf.raw))
            _lib.call("bl_dist_comm_connect_ipc", self.handle, b"".join(handles))
            g.barrier()  # every mailbox is zeroed and mapped before the first store

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
        if self._multi_process and self.world > 1:
            g = _group(self.group)
            handles = g.allgather_bytes(bytes(buf.raw))
            self._lib.call("bl_dist_comm_window_connect_ipc", self.handle, b"".join(handles))
            g.barrier()

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

    `comm=PeerComm(...)`: native route, the cross-rank sum rides in the streaming kernel's own last block over
    peer memory (`bl_dist_comm_activate`).  Otherwise the library's NCCL hook (`bl_dist_nccl_reduce_hook`): an
    `ncclAllReduce` of the reduced values enqueued on the library's stream between a kernel's local reduction and
    its epilogue.  Neither synchronises the host or calls back into Python."""

    def __init__(self, group=None, comm=None):
        from experiments_lanczos_adjoints_b200 import _lib

        self.group = group
        self.comm = comm
        self._lib = _lib
        self.error = None

    def __enter__(self):
        if self.comm is not None:
            self._lib.call("bl_dist_comm_activate", self.comm.handle)
        else:
            _group(self.group).install_reduce_hook(True)
        return self

    def __exit__(self, *exc):
        if self.comm is not None:
            self._lib.call("bl_dist_comm_activate", None)
        else:
            _group(self.group).install_reduce_hook(False)
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

        g = _group(group)
        self.group = group
        self.comm = comm
        self.rank = comm.rank if comm is not None else g.rank
        self.world = comm.world if comm is not None else g.world
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
        # one rank: a device copy; otherwise ncclAllGather on the library's stream
        _group(self.group).allgather_device(x_loc.ptr, full.ptr, x_loc.size, x_loc.dtype, dev.default_stream())
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

        hg = _group(group)
        self.group = group
        self.comm = comm
        self.rank = comm.rank if comm is not None else hg.rank
        self.world = comm.world if comm is not None else hg.world
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
        g, gs, item = self.g, self.gs, np.dtype(dtype).itemsize
        hg, s = _group(self.group), dev.default_stream()
        up = self.rank - 1 if self.has_top else None
        down = self.rank + 1 if self.has_bot else None
        last = field_ptr + (gs - 1) * g * item
        # two grouped ncclSend/ncclRecv pairs: first rows travel up while last rows arrive from above, then the
        # other direction (every rank makes the same two calls, so the pairs match up along the chain)
        hg.sendrecv_device(field_ptr, up, self.local.halo_ptr(slot_bot) if down is not None else None, down, g, dtype, s)
        hg.sendrecv_device(last, down, self.local.halo_ptr(slot_top) if up is not None else None, up, g, dtype, s)

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
