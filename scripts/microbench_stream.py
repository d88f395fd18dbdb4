"""Micro-benchmark of the basis-streaming kernels through the C ABI (bl_rows_dot / bl_rows_combine):
effective GB/s vs number of rows, repeated back to back (so small row counts run from L2)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import experiments_lanczos_adjoints_b200 as bl
from experiments_lanczos_adjoints_b200 import _lib
from experiments_lanczos_adjoints_b200 import device as dev

n = int(os.environ.get("N", 1_000_000))
K = 128
dtype = np.float32
ld = dev.basis_ld(n, dtype)
M = dev.zeros((K, n), dtype, ld=ld)
x = dev.asarray(np.random.default_rng(0).standard_normal(n).astype(dtype))
out = dev.empty((K,), dtype)
y = dev.empty((n,), dtype)
nbytes = _lib.load().bl_vec_workspace_bytes()
ws = dev.DeviceArray(((nbytes + 3) // 4,), np.float32)
s = dev.default_stream()
coef = np.ones(K)
reps = 20
for m in (1, 2, 4, 8, 12, 16, 24, 32, 48, 64, 96, 128):
    res = {}
    for name in ("dots", "combine"):
        def run():
            if name == "dots":
                _lib.call("bl_rows_dot", 0, n, m, M.ptr, ld, x.ptr, out.ptr, ws.ptr, nbytes, s.ptr)
            else:
                _lib.call("bl_rows_combine", 0, n, m, M.ptr, ld, coef.ctypes.data, 0, y.ptr, ws.ptr, nbytes, s.ptr)
        for _ in range(3):
            run()
        e0, e1 = bl.Event(), bl.Event()
        e0.record(s)
        for _ in range(reps):
            run()
        e1.record(s)
        e1.synchronize()
        us = e0.elapsed_ms(e1) / reps * 1e3
        res[name] = (us, (m + 1) * n * 4 / us / 1e3)
    print(f"m={m:4d}  dots {res['dots'][0]:8.1f} us {res['dots'][1]:7.0f} GB/s   combine {res['combine'][0]:8.1f} us {res['combine'][1]:7.0f} GB/s")
