from jax.scipy import linalg, signal  # noqa: F401
