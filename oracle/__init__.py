"""CPU oracle for the Lanczos/Arnoldi-adjoint hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`experiments_lanczos_adjoints_b200/`) may import this package; only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` do, and only as the checker or the timed CPU baseline.

The oracle is a NumPy restatement of the reference's algorithm
(`/root/reference/src/matfree_extensions/{arnoldi,lanczos,hutchinson}.py` and
the matvec backends under `util/`); each function cites the reference lines it
follows.

Pinning: the reference needs JAX, which is not installed in this image.  The
oracle is pinned against golden vectors under `tests/golden/` that were
produced by executing the UNMODIFIED reference sources
(`/root/reference/src/matfree_extensions/*.py`) on top of a small
torch-backed stand-in for the `jax` API (`oracle/jaxshim/`, generating script
`oracle/make_golden.py`).  The algorithm that produced the goldens is
therefore the reference's own code, line by line; what is NOT pinned is XLA's
floating-point summation order and JAX's threefry PRNG stream (probe vectors
are always passed explicitly).
"""

from oracle import krylov, operators  # noqa: F401
