"""Debug aid: the fused operator phase of k_step_tma against the separate launches, row by row."""
import os
import subprocess
import sys

import numpy as np

sys.path.insert(0, ".")
MODES = {"fused": {"BL_STEP_OP": "1"}, "step": {"BL_STEP_OP": "0", "BL_STEP": "2"}, "classic": {"BL_STEP": "0"}}

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import experiments_lanczos_adjoints_b200 as bl
    sys.path.insert(0, "tests")
    from test_gpu_parity import banded_spd

    n, K = int(sys.argv[3]), int(sys.argv[4])
    dtype = np.float32 if sys.argv[5] == "f32" else np.float64
    row, col, data = banded_spd(n, 4, seed=n)
    rng = np.random.default_rng(n + K)
    v = rng.standard_normal(n) + 2.0
    op = bl.operators.SparseOperator(row, col, (n, n))
    alg = bl.lanczos.tridiag(op, K, reortho="full")
    outs = []
    for rep in range(2):
        ((Qt, (alpha, beta)), (q_rem, b_rem)), pull = bl.vjp(alg, v.astype(dtype), data.astype(dtype))
        dalpha, dbeta = np.ones(K), np.ones(K - 1)
        dv, dp = pull(((None, (dalpha, dbeta)), (None, None)))
        outs.append((np.asarray(Qt.numpy()), np.asarray(alpha), np.asarray(beta), dv.numpy(), np.asarray(dp.numpy())))
    np.savez(sys.argv[2], Q0=outs[0][0], a0=outs[0][1], b0=outs[0][2], dv0=outs[0][3], dp0=outs[0][4],
             Q1=outs[1][0], a1=outs[1][1], b1=outs[1][2], dv1=outs[1][3], dp1=outs[1][4])
    sys.exit(0)

n, K, dt = (sys.argv[1:4] + ["20000", "6", "f32"][len(sys.argv) - 1:])[:3]
res = {}
for name, env in MODES.items():
    path = f"/tmp/dbg_{name}.npz"
    e = dict(os.environ)
    e.update(env)
    subprocess.run([sys.executable, __file__, "child", path, n, K, dt], check=True, env=e)
    res[name] = np.load(path)
ref = res["classic"]
for name in ("fused", "step"):
    r = res[name]
    print(f"== {name} vs classic (n={n}, K={K}, {dt})")
    print(" repeatable:", all(np.array_equal(r[k + "0"], r[k + "1"]) for k in ("Q", "a", "b", "dv", "dp")))
    print(" alpha", r["a0"], "\n  ref ", ref["a0"])
    print(" beta ", r["b0"], "\n  ref ", ref["b0"])
    for i in range(int(K)):
        d = r["Q0"][i] - ref["Q0"][i]
        bad = np.flatnonzero(np.abs(d) > 1e-4 * np.abs(ref["Q0"][i]).max())
        print(f" row {i}: |dQ|/|Q| = {np.linalg.norm(d) / np.linalg.norm(ref['Q0'][i]):.2e}; entries off: {bad.size}", bad[:8], bad[-4:] if bad.size else "")
    w = r["Q0"][1] * r["b0"][0] - ref["Q0"][1] * ref["b0"][0]  # difference of the unnormalised second basis vector
    bad = np.flatnonzero(np.abs(w) > 1e-4)
    print(" w'' entries off:", bad.size, "first", bad[:40], "values", w[bad[:8]], "q0 there", ref["Q0"][0][bad[:8]] * ref["a0"][0])
    if bad.size:
        runs = np.split(bad, np.flatnonzero(np.diff(bad) > 1) + 1)
        print(" contiguous runs:", [(int(x[0]), int(x[-1])) for x in runs[:30]], "count", len(runs))
    print(" dv err", np.linalg.norm(r["dv0"] - ref["dv0"]) / np.linalg.norm(ref["dv0"]), " dp err", np.linalg.norm(r["dp0"] - ref["dp0"]) / np.linalg.norm(ref["dp0"]))
