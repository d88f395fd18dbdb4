"""Host communicator of the multi-GPU layer: one process per GPU, no PyTorch.

What the multi-GPU paths need from the host side is small: a rendezvous (rank 0 hands out the 128-byte NCCL
unique id and the ranks swap 64-byte CUDA-IPC handles), a barrier, and sums of a few host numbers.  `SocketComm`
does that over one stream socket per rank to rank 0 -- a Unix-domain socket for the single-node case this
library targets (one 8 x B200 box), TCP when `MASTER_ADDR` is not local -- built from the torchrun environment
(`RANK`, `WORLD_SIZE`, `LOCAL_RANK`, `MASTER_ADDR`, `MASTER_PORT`).  Device buffers never travel through it:
`allreduce_device` / `allgather_device` / `sendrecv_device` are NCCL calls of libb200lanczos.so
(`bl_dist_nccl_*`, csrc/nccl_comm.cu) on the caller's stream -- "a single NCCL allreduce over NVLink" for the
probe-sharded estimator, the all-reduce / all-gather of the row-sharded operand.  Without a GPU (the CPU tests)
only the host collectives exist, which is all the sharding logic needs.
"""

from __future__ import annotations

import ctypes as C
import hashlib
import os
import socket
import struct
import time

import numpy as np

_DEFAULT = None  # the process-wide communicator set by `init_from_env`


class Comm:
    """Interface; `SelfComm` is the one-rank implementation."""

    rank, world, local_rank = 0, 1, 0

    # ---- host collectives ----
    def allgather_bytes(self, payload: bytes) -> list[bytes]:
        return [bytes(payload)]

    def allreduce_host(self, array, op: str = "sum") -> np.ndarray:
        return np.array(array, dtype=np.float64, copy=True)

    def barrier(self) -> None:
        return None

    def broadcast_bytes(self, payload: bytes | None, root: int = 0) -> bytes:
        return self.allgather_bytes(payload if self.rank == root else b"")[root]

    # ---- device collectives (NCCL; in place, on `stream`) ----
    def allreduce_device(self, ptr, count, dtype, stream, op: str = "sum") -> None:
        return None

    def allgather_device(self, send_ptr, recv_ptr, count, dtype, stream) -> None:
        from experiments_lanczos_adjoints_b200 import _lib

        _lib.call("bl_memcpy_d2d", recv_ptr, send_ptr, int(count) * np.dtype(dtype).itemsize, stream.ptr)

    def sendrecv_device(self, send_ptr, send_peer, recv_ptr, recv_peer, count, dtype, stream) -> None:
        return None

    def install_reduce_hook(self, on: bool = True) -> None:
        """Row sharding: every reduction of the Krylov loops on this thread is summed over the ranks."""
        return None

    def close(self) -> None:
        return None


class SelfComm(Comm):
    pass


def _recv_exact(sock, nbytes: int) -> bytes:
    chunks, got = [], 0
    while got < nbytes:
        chunk = sock.recv(min(nbytes - got, 1 << 22))
        if not chunk:
            raise ConnectionError("peer closed the rendezvous socket")
        chunks.append(chunk)
        got += len(chunk)
    return b"".join(chunks)


def _send_msg(sock, payload: bytes) -> None:
    sock.sendall(struct.pack("<Q", len(payload)) + payload)


def _recv_msg(sock) -> bytes:
    (nbytes,) = struct.unpack("<Q", _recv_exact(sock, 8))
    return _recv_exact(sock, nbytes)


class SocketComm(Comm):
    """Star topology over stream sockets: rank 0 listens, every other rank keeps one connection to it.  All
    collectives are "gather at rank 0, combine in rank order, send back" -- deterministic, and plenty for the few
    small host messages of the multi-GPU paths."""

    def __init__(self, rank: int, world: int, address, *, local_rank: int | None = None, timeout: float = 180.0,
                 token: bytes = b""):
        """`address`: a filesystem path (Unix-domain socket) or a `(host, port)` pair (TCP)."""
        self.rank, self.world = int(rank), int(world)
        self.local_rank = self.rank if local_rank is None else int(local_rank)
        self._nccl = None
        self._peers, self._sock, self._listener = [], None, None
        self._address = address
        family = socket.AF_UNIX if isinstance(address, str) else socket.AF_INET
        hello = b"b200lanczos-rdzv" + hashlib.sha256(token + str(world).encode()).digest()[:8]
        if self.world == 1:
            return
        if self.rank == 0:
            self._listener = socket.socket(family, socket.SOCK_STREAM)
            if family == socket.AF_UNIX:
                try:
                    os.unlink(address)
                except FileNotFoundError:
                    pass
            else:
                self._listener.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            self._listener.bind(address)
            self._listener.listen(self.world)
            self._listener.settimeout(timeout)
            peers = {}
            while len(peers) < self.world - 1:
                conn, _ = self._listener.accept()
                conn.settimeout(timeout)
                try:
                    msg = _recv_msg(conn)
                except (ConnectionError, socket.timeout, struct.error):
                    conn.close()
                    continue
                if not msg.startswith(hello):
                    conn.close()  # not one of ours
                    continue
                (peer_rank,) = struct.unpack("<I", msg[len(hello) :])
                peers[peer_rank] = conn
                _send_msg(conn, hello)
            self._peers = [peers[r] for r in range(1, self.world)]
        else:
            deadline = time.monotonic() + timeout
            while True:
                sock = socket.socket(family, socket.SOCK_STREAM)
                try:
                    sock.settimeout(5.0)
                    sock.connect(address)
                    _send_msg(sock, hello + struct.pack("<I", self.rank))
                    if _recv_msg(sock) == hello:
                        sock.settimeout(timeout)
                        self._sock = sock
                        break
                except (OSError, ConnectionError, struct.error):
                    pass
                sock.close()
                if time.monotonic() > deadline:
                    raise TimeoutError(f"rank {self.rank}: no rendezvous at {address!r}")
                time.sleep(0.05)

    # ---- host collectives ----
    def allgather_bytes(self, payload: bytes) -> list[bytes]:
        payload = bytes(payload)
        if self.world == 1:
            return [payload]
        if self.rank == 0:
            parts = [payload] + [_recv_msg(c) for c in self._peers]
            blob = b"".join(struct.pack("<Q", len(p)) + p for p in parts)
            for c in self._peers:
                _send_msg(c, blob)
            return parts
        _send_msg(self._sock, payload)
        blob, parts, pos = _recv_msg(self._sock), [], 0
        while pos < len(blob):
            (nbytes,) = struct.unpack_from("<Q", blob, pos)
            parts.append(blob[pos + 8 : pos + 8 + nbytes])
            pos += 8 + nbytes
        return parts

    def allreduce_host(self, array, op: str = "sum") -> np.ndarray:
        shape = np.shape(array)
        arr = np.array(array, dtype=np.float64, copy=True).reshape(-1)
        if self.world == 1:
            return arr.reshape(shape)
        if self.rank == 0:
            total = arr
            for c in self._peers:  # rank order: the same result on every run
                other = np.frombuffer(_recv_msg(c), dtype=np.float64)
                total = np.maximum(total, other) if op == "max" else total + other
            blob = total.tobytes()
            for c in self._peers:
                _send_msg(c, blob)
            return total.reshape(shape)
        _send_msg(self._sock, arr.tobytes())
        return np.frombuffer(_recv_msg(self._sock), dtype=np.float64).copy().reshape(shape)

    def barrier(self) -> None:
        self.allgather_bytes(b"")

    # ---- NCCL ----
    def nccl(self):
        """The NCCL communicator of this group (created on first use, on the current device)."""
        if self._nccl is None:
            from experiments_lanczos_adjoints_b200 import _lib

            ident = C.create_string_buffer(128)
            if self.rank == 0:
                _lib.call("bl_dist_nccl_unique_id", ident)
            raw = self.broadcast_bytes(bytes(ident.raw), root=0)
            handle = C.c_void_p()
            _lib.call("bl_dist_nccl_init", raw, self.rank, self.world, C.byref(handle))
            self._nccl = handle.value
        return self._nccl

    def allreduce_device(self, ptr, count, dtype, stream, op: str = "sum") -> None:
        if self.world == 1 or int(count) == 0:
            return
        from experiments_lanczos_adjoints_b200 import _lib, device as dev

        _lib.call("bl_dist_nccl_allreduce", self.nccl(), ptr, int(count), dev.dtype_code(dtype), 1 if op == "max" else 0,
                  stream.ptr)  # fmt: skip

    def allgather_device(self, send_ptr, recv_ptr, count, dtype, stream) -> None:
        if self.world == 1:
            return super().allgather_device(send_ptr, recv_ptr, count, dtype, stream)
        from experiments_lanczos_adjoints_b200 import _lib, device as dev

        _lib.call("bl_dist_nccl_allgather", self.nccl(), send_ptr, recv_ptr, int(count), dev.dtype_code(dtype), stream.ptr)

    def sendrecv_device(self, send_ptr, send_peer, recv_ptr, recv_peer, count, dtype, stream) -> None:
        if self.world == 1:
            return
        from experiments_lanczos_adjoints_b200 import _lib, device as dev

        _lib.call("bl_dist_nccl_sendrecv", self.nccl(), send_ptr, -1 if send_peer is None else int(send_peer), recv_ptr,
                  -1 if recv_peer is None else int(recv_peer), int(count), dev.dtype_code(dtype), stream.ptr)  # fmt: skip

    def install_reduce_hook(self, on: bool = True) -> None:
        from experiments_lanczos_adjoints_b200 import _lib

        _lib.call("bl_dist_nccl_reduce_hook", self.nccl() if (on and self.world > 1) else None)

    def close(self) -> None:
        if self._nccl is not None:
            from experiments_lanczos_adjoints_b200 import _lib

            try:
                _lib.call("bl_dist_nccl_destroy", self._nccl)
            finally:
                self._nccl = None
        for c in self._peers:
            c.close()
        self._peers = []
        if self._sock is not None:
            self._sock.close()
            self._sock = None
        if self._listener is not None:
            self._listener.close()
            self._listener = None
            if isinstance(self._address, str):
                try:
                    os.unlink(self._address)
                except OSError:
                    pass


def rendezvous_address(master_addr: str, master_port: int):
    """Where rank 0 listens.  torchrun's own store owns `MASTER_PORT`, so: a Unix-domain socket named after the
    port and the launcher's pid (all workers of one `torchrun` share the parent; a new launch is a new name) when
    the master is this machine, else TCP on `MASTER_PORT + 1` (`BL_RDZV_PORT` overrides)."""
    override = os.environ.get("BL_RDZV_PORT")
    if override:
        return (master_addr, int(override))
    if master_addr in ("127.0.0.1", "localhost", "::1", socket.gethostname()):
        return f"/tmp/bl_rdzv_{int(master_port)}_{os.getppid()}.sock"
    return (master_addr, int(master_port) + 1)


def init_from_env() -> Comm:
    """The process-wide communicator from the torchrun environment; binds this process to GPU `LOCAL_RANK` when a
    device is present.  A single-process run gets `SelfComm`.  Idempotent."""
    global _DEFAULT
    if _DEFAULT is not None:
        return _DEFAULT
    from experiments_lanczos_adjoints_b200 import device as dev

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if dev.device_count() > 0:
        dev.set_device(local_rank)
    if world == 1:
        _DEFAULT = SelfComm()
        return _DEFAULT
    addr = os.environ.get("MASTER_ADDR", "127.0.0.1")
    port = int(os.environ.get("MASTER_PORT", 29500))
    token = f"{addr}:{port}:{os.environ.get('TORCHELASTIC_RUN_ID', '')}".encode()
    _DEFAULT = SocketComm(rank, world, rendezvous_address(addr, port), local_rank=local_rank, token=token)
    return _DEFAULT


def default() -> Comm:
    """The communicator `init_from_env` made, or the one-rank communicator."""
    return _DEFAULT if _DEFAULT is not None else SelfComm()


def shutdown() -> None:
    global _DEFAULT
    if _DEFAULT is not None:
        _DEFAULT.close()
        _DEFAULT = None
