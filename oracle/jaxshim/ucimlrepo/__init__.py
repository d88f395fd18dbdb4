"""Empty stub so `util/exp_util.py` imports; the UCI dataset download is out of scope (no network)."""
