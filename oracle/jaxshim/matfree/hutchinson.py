import jax
import jax.numpy as jnp


def hutchinson(integrand_fun, /, sample_fun):
    def sample(key, *parameters):
        samples = sample_fun(key)
        Qs = jax.vmap(lambda vec: integrand_fun(vec, *parameters))(samples)
        return jax.tree_util.tree_map(lambda s: jnp.mean(s, axis=0), Qs)

    return sample
