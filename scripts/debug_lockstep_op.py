"""Debug aid: lockstep batches (BatchedTridiagAdjointPlan) with the operator call inside the step kernel against
BL_STEP_OP=0, one lane and two lanes in flight, at a given size."""
import os
import subprocess
import sys

import numpy as np

sys.path.insert(0, ".")

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import experiments_lanczos_adjoints_b200 as bl
    from experiments_lanczos_adjoints_b200 import device as dev, plan as bl_plan, synthetic

    n, K, P, lanes = (int(a) for a in sys.argv[3:7])
    dtype = np.float32 if sys.argv[7] == "f32" else np.float64
    row, col, data = synthetic.banded_spd_coo(n, 5, seed=0)
    rng = np.random.default_rng(1)
    ops = [bl.operators.SparseOperator(row, col, (n, n))]
    ops += [ops[0].clone() for _ in range(lanes - 1)]
    plans = [bl_plan.BatchedTridiagAdjointPlan(o, K, dtype, P, stream=dev.Stream()) for o in ops]
    dH = np.stack([synthetic.slq_cotangent_dH(rng.standard_normal(K), rng.standard_normal(K - 1), dtype) for _ in range(P)])
    vs = [(rng.integers(0, 2, size=(P, n)) * 2 - 1).astype(dtype) / np.sqrt(n) for _ in range(lanes)]
    out = {}
    with dev.blocks_per_sm(int(os.environ.get('DBG_BPS', 1 if lanes >= 2 else dev.get_blocks_per_sm()))):
        for rep in range(int(os.environ.get('DBG_REPS', 2))):
            for pl, v in zip(plans, vs):
                pl.set_vectors(v)
                pl.set_params(data.astype(dtype))
                pl.set_cotangents(dH)
            dev.synchronize()
            for pl in plans:
                pl.forward()
            for pl in plans:
                pl.adjoint()
            dev.synchronize()
            for li, pl in enumerate(plans):
                out[f"H{li}_{rep}"] = pl.H.numpy(pl.stream)
                out[f"dv{li}_{rep}"] = pl.dv.numpy(pl.stream)
                out[f"g{li}_{rep}"] = pl.grads[0].numpy(pl.stream)
                L = pl.Lam.numpy(pl.stream).reshape(P, K, -1)
                if np.isnan(L).any():
                    for p_ in range(P):
                        nanrows = [int(k_) for k_ in range(K) if np.isnan(L[p_, k_]).any()]
                        if nanrows:
                            k0_ = max(nanrows)
                            idxs = np.flatnonzero(np.isnan(L[p_, k0_]))
                            print(f"rep {rep} lane {li} run {p_}: Lambda rows with NaN {nanrows}; in row {k0_}: {idxs.size} entries, "
                                  f"first {idxs[:6]}, last {idxs[-3:]}; contiguous runs {np.flatnonzero(np.diff(idxs) > 1).size + 1}", flush=True)
    reps = int(os.environ.get('DBG_REPS', 2))
    for li in range(lanes):
        bad = [rep for rep in range(reps) if not all(np.array_equal(out[f"{k}{li}_{rep}"], out[f"{k}{li}_0"]) for k in ("H", "dv", "g"))]
        print(f"lane {li}: reps that differ from rep 0: {bad}", flush=True)
    np.savez(sys.argv[2], **{k: v for k, v in out.items() if int(k.split("_")[1]) < 3})
    sys.exit(0)

n, K, P, lanes, dt = (sys.argv[1:6] + ["1000000", "12", "4", "2", "f32"][len(sys.argv) - 1:])[:5]
res = {}
for name, env in {"fused": {"BL_STEP_OP": "1"}, "separate": {"BL_STEP_OP": "0"}}.items():
    path = f"/tmp/dbgl_{name}.npz"
    e = dict(os.environ)
    e.update(env)
    subprocess.run([sys.executable, __file__, "child", path, n, K, P, lanes, dt], check=True, env=e)
    res[name] = np.load(path)
f, s = res["fused"], res["separate"]
for k in sorted(s.files, key=lambda k: (k.split("_")[1], k)):
    a, b = f[k].astype(np.float64), s[k].astype(np.float64)
    print(f"{k:8s} nan {int(np.isnan(a).sum()):8d} (separate {int(np.isnan(b).sum())})  rel diff "
          f"{np.linalg.norm(np.nan_to_num(a) - np.nan_to_num(b)) / max(np.linalg.norm(np.nan_to_num(b)), 1e-300):.3e}")
for k in sorted(f.files):
    if k.startswith("H") and np.isnan(f[k]).any():
        Hn = np.isnan(f[k].reshape(int(P), int(K), int(K)))
        print(k, "first NaN column per run:", [int(np.argmax(Hn[p].any(axis=0))) if Hn[p].any() else -1 for p in range(int(P))])
        Hp = f[k].reshape(int(P), int(K), int(K)); Hs = s[k].reshape(int(P), int(K), int(K))
        p = int(np.argmax([Hn[q].any() for q in range(int(P))])); c = int(np.argmax(Hn[p].any(axis=0)))
        print("  run", p, "columns", max(0, c - 2), "..", c, "fused:\n", Hp[p][:, max(0, c - 2):c + 1].T, "\n  separate:\n", Hs[p][:, max(0, c - 2):c + 1].T)
H = f["H0_0"].reshape(int(P), int(K), int(K))
print("alpha run 0:", np.diag(H[0])[:6], " separate:", np.diag(s["H0_0"].reshape(int(P), int(K), int(K))[0])[:6])
