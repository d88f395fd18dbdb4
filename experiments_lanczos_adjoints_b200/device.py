"""Device arrays, streams and events on top of the C ABI (no torch, no cupy).

`DeviceArray` is the host layer's stand-in for a `jax.Array` living on the GPU: an owned
(or borrowed) device pointer with a shape and a dtype.  2-D arrays are row-major with a
row stride `ld` (in elements); a Krylov basis is stored as `K` rows of length `ld >= n`
(the reference's `Q.T`) and shown to the caller as `(n, K)` through a zero-copy `.T` view.
"""

from __future__ import annotations

import ctypes as C
import threading
import weakref

import numpy as np

from experiments_lanczos_adjoints_b200 import _lib

_DTYPES = {np.dtype(np.float32): _lib.BL_F32, np.dtype(np.float64): _lib.BL_F64}


def dtype_code(dtype) -> int:
    try:
        return _DTYPES[np.dtype(dtype)]
    except KeyError:
        raise TypeError(f"only float32 and float64 run on the device, got {dtype}") from None


def device_count() -> int:
    n = C.c_int(0)
    _lib.call("bl_device_count", C.byref(n))
    return n.value


def set_device(index: int):
    _lib.call("bl_set_device", int(index))


def sm_count() -> int:
    n = C.c_int(0)
    _lib.call("bl_device_sm_count", C.byref(n))
    return n.value


def set_blocks_per_sm(blocks: int):
    """Blocks per SM of the basis-streaming kernels from now on: 2 (default), 1 (several runs in flight on separate
    streams), 0 = back to the default (`bl_set_blocks_per_sm`)."""
    _lib.call("bl_set_blocks_per_sm", int(blocks))


def get_blocks_per_sm() -> int:
    b = C.c_int(0)
    _lib.call("bl_get_blocks_per_sm", C.byref(b))
    return b.value


class blocks_per_sm:
    """`with blocks_per_sm(1): ...` -- set the process-wide value for a region and put the caller's value back."""

    def __init__(self, blocks: int):
        self.blocks = int(blocks)

    def __enter__(self):
        self.saved = get_blocks_per_sm()
        set_blocks_per_sm(self.blocks)
        return self

    def __exit__(self, *exc):
        set_blocks_per_sm(self.saved)
        return False


def launch_count() -> int:
    n = C.c_uint64(0)
    _lib.call("bl_launch_count", C.byref(n))
    return n.value


class Stream:
    def __init__(self):
        p = C.c_void_p()
        _lib.call("bl_stream_create", C.byref(p))
        self.ptr = p.value
        self._fin = weakref.finalize(self, _destroy_stream, self.ptr)
        _thread_streams().add(self)  # the streams THIS host thread enqueues on (see _Pool)

    def synchronize(self):
        _lib.call("bl_stream_sync", self.ptr)

    def wait_event(self, event):
        """Work enqueued on this stream afterwards starts after `event` (recorded on another stream)."""
        _lib.call("bl_stream_wait_event", self.ptr, event.ptr)


def _destroy_stream(ptr):
    try:
        _lib.load().bl_stream_destroy(ptr)
    except Exception:
        pass


_tls = threading.local()


def _thread_streams():
    streams = getattr(_tls, "streams", None)
    if streams is None:
        streams = _tls.streams = weakref.WeakSet()
    return streams


def default_stream() -> Stream:
    """The stream every call without an explicit `stream=` uses: one per HOST THREAD, so several
    threads can drive independent work (the C library keeps no global stream either)."""
    st = getattr(_tls, "stream", None)
    if st is None:
        st = _tls.stream = Stream()
    return st


def synchronize():
    _lib.call("bl_device_sync")


class Event:
    def __init__(self):
        p = C.c_void_p()
        _lib.call("bl_event_create", C.byref(p))
        self.ptr = p.value
        self._fin = weakref.finalize(self, _destroy_event, self.ptr)

    def record(self, stream: Stream | None = None):
        _lib.call("bl_event_record", self.ptr, (stream or default_stream()).ptr)

    def synchronize(self):
        _lib.call("bl_event_sync", self.ptr)

    def elapsed_ms(self, end: "Event") -> float:
        ms = C.c_float(0)
        _lib.call("bl_event_elapsed_ms", self.ptr, end.ptr, C.byref(ms))
        return ms.value


def _destroy_event(ptr):
    try:
        _lib.load().bl_event_destroy(ptr)
    except Exception:
        pass


# --- a small caching allocator: cudaMalloc/cudaFree synchronise, the loops must not -----
class _Pool:
    """One pool per HOST THREAD: a returned buffer is reused without synchronisation, which is safe for work on
    ONE stream (stream order) -- and every thread has its own default stream.  Once the thread has made further
    streams (`stream=` arguments, plans, lanes), a buffer freed while work on stream A still uses it could be handed
    to work on stream B: buffers returned since the last synchronisation are therefore marked, and reusing a marked
    buffer synchronises the thread's streams first (only in programs that use several streams; the Krylov loops
    themselves never allocate).  A buffer finalised from another thread (garbage collection) still goes back to
    its owner's pool."""

    def __init__(self, max_cached_bytes=24 << 30):
        self.free = {}
        self.cached = 0
        self.max_cached = max_cached_bytes
        self.lock = threading.Lock()
        self.in_doubt = set()  # returned since the last synchronisation

    def alloc(self, nbytes: int) -> tuple[int, int]:
        size = max(256, (int(nbytes) + 255) // 256 * 256)
        ptr = None
        streams = list(_thread_streams())
        with self.lock:
            bucket = self.free.get(size)
            if bucket:
                self.cached -= size
                ptr = bucket.pop()
                if ptr not in self.in_doubt or len(streams) <= 1:
                    self.in_doubt.discard(ptr)
                    return ptr, size
        if ptr is not None:
            for st in streams:  # this thread's streams are idle now: whatever used its returned buffers has finished
                st.synchronize()
            with self.lock:
                self.in_doubt.clear()
            return ptr, size
        p = C.c_void_p()
        try:
            _lib.call("bl_malloc", C.byref(p), size)
        except MemoryError:
            self.release_all()
            _lib.call("bl_malloc", C.byref(p), size)
        return p.value, size

    def give_back(self, ptr: int, size: int):
        with self.lock:
            if self.cached + size <= self.max_cached:
                self.free.setdefault(size, []).append(ptr)
                self.cached += size
                self.in_doubt.add(ptr)
                return
        try:
            _lib.load().bl_free(ptr)
        except Exception:
            pass

    def release_all(self):
        with self.lock:
            buckets, self.free, self.cached = self.free, {}, 0
            self.in_doubt.clear()
        for bucket in buckets.values():
            for ptr in bucket:
                _lib.load().bl_free(ptr)


_pools = []
_pools_lock = threading.Lock()


def _thread_pool() -> _Pool:
    pool = getattr(_tls, "pool", None)
    if pool is None:
        pool = _tls.pool = _Pool()
        with _pools_lock:
            _pools.append(pool)
    return pool


def empty_cache():
    synchronize()
    with _pools_lock:
        pools = list(_pools)
    for pool in pools:
        pool.release_all()


class _Owner:
    """Owns one pool allocation; shared by every view of it."""

    def __init__(self, nbytes):
        pool = _thread_pool()
        self.ptr, self.size = pool.alloc(nbytes)
        self._fin = weakref.finalize(self, pool.give_back, self.ptr, self.size)


class DeviceArray:
    """Row-major device array (1-D, or 2-D with row stride `ld`); `.T` is a zero-copy view."""

    __array_priority__ = 100

    def __init__(self, shape, dtype, *, ld=None, owner=None, ptr=None, transposed=False):
        self.dtype = np.dtype(dtype)
        dtype_code(self.dtype)
        self._shape = tuple(int(s) for s in shape)  # storage shape (rows, cols) or (n,)
        if len(self._shape) > 2:
            raise ValueError("DeviceArray supports 0-, 1- and 2-D arrays")
        self.ld = int(ld) if ld is not None else (self._shape[-1] if self._shape else 1)
        self._transposed = bool(transposed)
        if owner is None and ptr is None:
            owner = _Owner(self.storage_elems * self.dtype.itemsize)
            ptr = owner.ptr
        self._owner = owner
        self.ptr = int(ptr) if ptr is not None else owner.ptr

    # -- shape bookkeeping --
    @property
    def storage_elems(self) -> int:
        if len(self._shape) == 2:
            return self._shape[0] * self.ld
        return int(np.prod(self._shape, dtype=np.int64)) if self._shape else 1

    @property
    def shape(self):
        return self._shape[::-1] if self._transposed else self._shape

    @property
    def ndim(self):
        return len(self._shape)

    @property
    def size(self):
        return int(np.prod(self._shape, dtype=np.int64)) if self._shape else 1

    @property
    def T(self):
        if self.ndim < 2:
            return self
        return DeviceArray(self._shape, self.dtype, ld=self.ld, owner=self._owner, ptr=self.ptr,
                           transposed=not self._transposed)  # fmt: skip

    @property
    def is_transposed(self):
        return self._transposed

    def row(self, j: int) -> "DeviceArray":
        """Zero-copy view of storage row `j` (a basis vector)."""
        rows, cols = self._shape
        if not 0 <= j < rows:
            raise IndexError(j)
        off = j * self.ld * self.dtype.itemsize
        return DeviceArray((cols,), self.dtype, owner=self._owner, ptr=self.ptr + off)

    def __len__(self):
        return self.shape[0]

    # -- transfers --
    def numpy(self, stream: Stream | None = None) -> np.ndarray:
        stream = stream or default_stream()
        host = np.empty(self.storage_elems, dtype=self.dtype)
        _lib.call("bl_memcpy_d2h", host.ctypes.data, self.ptr, host.nbytes, stream.ptr)
        stream.synchronize()
        if self.ndim == 2:
            rows, cols = self._shape
            out = host.reshape(rows, self.ld)[:, :cols]
            return np.ascontiguousarray(out.T if self._transposed else out)
        return host.reshape(self._shape)

    def __array__(self, dtype=None, copy=None):
        out = self.numpy()
        return out.astype(dtype) if dtype is not None else out

    def __float__(self):
        return float(self.numpy().reshape(-1)[0])

    def copy(self, stream: Stream | None = None) -> "DeviceArray":
        out = DeviceArray(self._shape, self.dtype, ld=self.ld, transposed=self._transposed)
        _lib.call("bl_memcpy_d2d", out.ptr, self.ptr, self.storage_elems * self.dtype.itemsize,
                  (stream or default_stream()).ptr)  # fmt: skip
        return out

    def __repr__(self):
        return f"DeviceArray(shape={self.shape}, dtype={self.dtype}, ld={self.ld})"

    @property
    def __cuda_array_interface__(self):
        if self.ndim == 2:
            strides = (self.ld * self.dtype.itemsize, self.dtype.itemsize)
            if self._transposed:
                strides = strides[::-1]
        else:
            strides = None
        return {"shape": self.shape, "typestr": self.dtype.str, "data": (self.ptr, False),
                "version": 3, "strides": strides}  # fmt: skip


def basis_ld(n: int, dtype) -> int:
    """Row stride of a basis buffer: n rounded up to 128 bytes (16-byte rule of the C ABI)."""
    per = 128 // np.dtype(dtype).itemsize
    return (int(n) + per - 1) // per * per


def empty(shape, dtype, ld=None) -> DeviceArray:
    if isinstance(shape, (int, np.integer)):
        shape = (int(shape),)
    return DeviceArray(shape, dtype, ld=ld)


def zeros(shape, dtype, ld=None, stream: Stream | None = None) -> DeviceArray:
    out = empty(shape, dtype, ld=ld)
    _lib.call("bl_memset", out.ptr, 0, out.storage_elems * out.dtype.itemsize, (stream or default_stream()).ptr)
    return out


def asarray(x, dtype=None, stream: Stream | None = None) -> DeviceArray:
    """Host -> device (a `DeviceArray` of the right dtype is passed through)."""
    if isinstance(x, DeviceArray):
        if dtype is not None and np.dtype(dtype) != x.dtype:
            return asarray(x.numpy().astype(dtype), stream=stream)
        return x
    host = np.ascontiguousarray(np.asarray(x, dtype=dtype))
    if host.dtype not in _DTYPES:
        host = host.astype(np.float64 if host.dtype.itemsize > 4 else np.float32)
    out = DeviceArray(host.shape, host.dtype)
    stream = stream or default_stream()
    _lib.call("bl_memcpy_h2d", out.ptr, host.ctypes.data, host.nbytes, stream.ptr)
    stream.synchronize()  # `host` may be a temporary
    return out


def basis_from_host(mat_kn: np.ndarray, dtype) -> DeviceArray:
    """Upload a `(K, n)` host matrix into basis layout (K rows of stride `basis_ld(n)`)."""
    mat_kn = np.asarray(mat_kn, dtype=dtype)
    K, n = mat_kn.shape
    ld = basis_ld(n, dtype)
    host = np.zeros((K, ld), dtype=dtype)
    host[:, :n] = mat_kn
    out = DeviceArray((K, n), dtype, ld=ld)
    stream = default_stream()
    _lib.call("bl_memcpy_h2d", out.ptr, host.ctypes.data, host.nbytes, stream.ptr)
    stream.synchronize()
    return out
