import numpy as _np


def PRNGKey(seed):
    return _np.array([0, seed], dtype=_np.uint32)


def uniform(key, shape=(), dtype=float, minval=0.0, maxval=1.0):
    rng = _np.random.default_rng(int(key[-1]))
    return rng.uniform(minval, maxval, size=shape)
