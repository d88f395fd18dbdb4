from jax._core import Partial, tree_flatten, tree_map, tree_unflatten  # noqa: F401
