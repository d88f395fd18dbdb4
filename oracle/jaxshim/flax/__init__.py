"""Empty stub so `util/pde_util.py` imports; the flax MLP is out of scope."""
