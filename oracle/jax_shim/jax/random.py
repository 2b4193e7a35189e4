"""PRNG is deliberately absent: JAX's threefry stream cannot be reproduced."""


def _na(*a, **k):
    raise NotImplementedError("jax.random is not provided by the oracle shim")


key = split = uniform = multivariate_normal = normal = _na
