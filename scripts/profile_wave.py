import sys, os, json
import numpy as np
sys.path.insert(0, "/root/repo")
import experiments_lanczos_adjoints_b200 as bl
from experiments_lanczos_adjoints_b200 import plan as bl_plan
g = 4096
dx = 1.0 / (g - 1)
op = bl.operators.WaveStencilOperator(g, bl.operators.WaveStencilOperator.stencil_laplacian(dx) * dx * dx)
xs = np.linspace(0, 1, g)
y0 = np.stack([np.exp(-80 * ((xs[:, None] - 0.4) ** 2 + (xs[None, :] - 0.6) ** 2)), np.zeros((g, g))]).astype(np.float32)
scale = (1.0 + 0.1 * np.sin(6 * xs)[:, None] * np.cos(4 * xs)[None, :]).astype(np.float32)
alg = bl.arnoldi.hessenberg(op, 10, reortho="full")
v = bl.asarray(y0.ravel()); sc = bl.asarray(scale)
def run():
    (Q, H, r, c), pull = bl.vjp(alg, v, sc)
    return pull((None, np.eye(10, dtype=np.float32), None, None))
run(); run()
prof = bl_plan.profile(run)
print(json.dumps({k: (c["launches"], round(c["ms"], 3), round(c["algorithmic_bytes"] / max(c["ms"], 1e-9) / 1e6)) for k, c in prof.items()}))
