"""Debug aid: where do two lockstep lanes in flight first deviate from the same lanes run one after the other?
Every lane runs alone first (reference), then all lanes together for DBG_REPS repetitions; per lane and run the
first basis row / Lambda row that differs is reported with the positions of the differing entries.

usage: debug_lockstep_diff.py n K P lanes f32|f64"""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
import experiments_lanczos_adjoints_b200 as bl  # noqa: E402
from experiments_lanczos_adjoints_b200 import device as dev, plan as bl_plan, synthetic  # noqa: E402

n, K, P, lanes = (int(a) for a in sys.argv[1:5])
dtype = np.float32 if sys.argv[5] == "f32" else np.float64
row, col, data = synthetic.banded_spd_coo(n, 5, seed=0)
rng = np.random.default_rng(1)
ops = [bl.operators.SparseOperator(row, col, (n, n))]
ops += [ops[0].clone() for _ in range(lanes - 1)]
plans = [bl_plan.BatchedTridiagAdjointPlan(o, K, dtype, P, stream=dev.Stream()) for o in ops]
dH = np.stack([synthetic.slq_cotangent_dH(rng.standard_normal(K), rng.standard_normal(K - 1), dtype) for _ in range(P)])
vs = [(rng.integers(0, 2, size=(P, n)) * 2 - 1).astype(dtype) / np.sqrt(n) for _ in range(lanes)]
G = 148 * int(os.environ.get("DBG_BPS", 1))


def describe(name, got, ref, descending):
    """got, ref: (P, K, n)"""
    for p in range(P):
        order = range(K - 1, -1, -1) if descending else range(K)
        for k in order:
            a, b = got[p, k], ref[p, k]
            same = (a == b) | (np.isnan(a) & np.isnan(b))
            if same.all():
                continue
            idx = np.flatnonzero(~same)
            breaks = np.flatnonzero(np.diff(idx) > 1)
            starts = np.concatenate([[idx[0]], idx[breaks + 1]])[:6]
            ends = np.concatenate([idx[breaks], [idx[-1]]])[:6]
            d = np.abs(a[idx].astype(np.float64) - b[idx].astype(np.float64))
            scale = np.abs(b).max()
            print(f"   {name} run {p}: first differing row {k}: {idx.size} entries in {breaks.size + 1} stretches "
                  f"{list(zip(starts.tolist(), ends.tolist()))}; max |diff| {np.nanmax(d) if np.isfinite(d).any() else float('nan'):.3e} "
                  f"(row max {scale:.3e}), NaN {int(np.isnan(a).sum())}; block of first entry at 1 block/SM: "
                  f"{idx[0] * 148 // n}", flush=True)
            break


def dbg_read(tag):
    import ctypes as C
    from experiments_lanczos_adjoints_b200 import _lib
    lib = _lib.load()
    if not hasattr(lib, "bl_step_debug_read"):
        return
    out = (C.c_ulonglong * 16)()
    lib.bl_step_debug_read.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
    lib.bl_step_debug_read(out, 1)
    if out[0]:
        f = lambda u: np.array([u], np.uint64).view(np.float64)[0]
        a = np.array([out[5]], np.uint64).view(np.int64)[0]
        print(f"   [{tag}] first bad value: code {out[0]} block {out[1]} thread {out[2]} run {out[3]} step {np.array([out[4]], np.uint64).view(np.int64)[0]} "
              f"a {a} v0 {f(out[6])} v1 {f(out[7])} sm {out[8]}", flush=True)
    else:
        print(f"   [{tag}] no bad value recorded", flush=True)


def setup():
    for pl, v in zip(plans, vs):
        pl.set_vectors(v)
        pl.set_params(data.astype(dtype))
        pl.set_cotangents(dH)
    dev.synchronize()


def fetch(pl, arr):
    return arr.numpy(pl.stream).reshape(P, K, -1)[:, :, :n]


with dev.blocks_per_sm(int(os.environ.get("DBG_BPS", 1))):
    setup()
    ref = []
    for pl in plans:  # one lane at a time
        pl.forward()
        dev.synchronize()
        Q = fetch(pl, pl.Q).copy()
        H = pl.H.numpy(pl.stream).copy()
        pl.adjoint()
        dev.synchronize()
        ref.append((Q, H, fetch(pl, pl.Lam).copy()))
    for rep in range(int(os.environ.get("DBG_REPS", 3))):
        setup()
        dbg_read("before")
        for pl in plans:
            pl.forward()
        dev.synchronize()
        dbg_read("forward")
        for li, pl in enumerate(plans):
            Q = fetch(pl, pl.Q)
            H = pl.H.numpy(pl.stream)
            okQ, okH = np.array_equal(Q, ref[li][0], equal_nan=True), np.array_equal(H, ref[li][1], equal_nan=True)
            print(f"rep {rep} lane {li} forward: Q {'same' if okQ else 'DIFFERS'}, H {'same' if okH else 'DIFFERS'}", flush=True)
            if not okQ:
                describe("Q", Q, ref[li][0], False)
            if not okH:
                Hd = (H != ref[li][1]).reshape(P, K, K)
                for p in range(P):
                    if Hd[p].any():
                        cols = np.flatnonzero(Hd[p].any(axis=0))
                        print(f"   H run {p}: first differing column {cols[0]}: rows {np.flatnonzero(Hd[p][:, cols[0]]).tolist()}, "
                              f"got {H.reshape(P, K, K)[p][Hd[p][:, cols[0]], cols[0]]}, "
                              f"ref {ref[li][1].reshape(P, K, K)[p][Hd[p][:, cols[0]], cols[0]]}", flush=True)
        if os.environ.get("DBG_RESTORE", "1") == "1":  # the adjoint starts from the reference forward of every lane
            for li, pl in enumerate(plans):
                pl.forward()
                dev.synchronize()
        dbg_read("restore")
        for pl in plans:
            pl.adjoint()
        dev.synchronize()
        dbg_read("adjoint")
        for li, pl in enumerate(plans):
            L = fetch(pl, pl.Lam)
            ok = np.array_equal(L, ref[li][2], equal_nan=True)
            print(f"rep {rep} lane {li} adjoint: Lambda {'same' if ok else 'DIFFERS'}", flush=True)
            if not ok:
                describe("Lambda", L, ref[li][2], True)
