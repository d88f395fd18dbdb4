"""Empty stub so `util/pde_util.py` imports; the diffrax solvers are out of scope."""
