"""NumPy-backed stand-in for the `jax` package -- TEST INFRASTRUCTURE ONLY.

Purpose: execute the *unmodified* reference sources under /root/reference/src
(tumaer/JAXFLUIDS, pure Python on jax.numpy) in fp64 on the CPU so that the
golden vectors under tests/golden/ come from the reference's own code, not a
restatement.  Only the surface that the single-phase convective path touches
is provided.  Nothing in the product package imports this.

Semantics that matter (JAX arrays are immutable):
  * `x += y` on an array must NOT alias (the reference relies on it, e.g.
    time_integration/time_integrator.py:52-57, halos/outer/material.py:892),
    so the ndarray subclass returns fresh arrays from in-place operators.
  * `x.at[idx].set/add/mul/min/max(v)` returns a modified copy.
"""
import numpy as _np
from . import numpy  # noqa: F401  (jax.numpy)
from . import lax, experimental, scipy, tree_util, random, nn, version, tree  # noqa: F401
from .numpy import ndarray as Array
from .numpy import ndarray  # jax.ndarray

__version__ = version.__version__


class _Config:
    def __init__(self):
        self._v = {"jax_enable_x64": True, "jax_disable_jit": False}

    def update(self, k, v):
        self._v[k] = v

    def read(self, k):
        return self._v.get(k)

    def __getattr__(self, k):
        try:
            return self.__dict__["_v"][k]
        except KeyError:
            raise AttributeError(k)


config = _Config()


def _identity_transform(fun=None, *a, **k):
    if fun is None or not callable(fun):
        return lambda f: f
    return fun


jit = _identity_transform
checkpoint = _identity_transform
remat = _identity_transform


def pmap(fun=None, *a, **k):
    # single-block oracle only: pmap'd functions are never called.
    return fun


def vmap(fun, in_axes=0, out_axes=0, **k):
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                n = _np.shape(a)[ax]
                break
        outs = []
        for i in range(n):
            sl = [(_np.take(a, i, axis=ax) if ax is not None else a) for a, ax in zip(args, axes)]
            outs.append(fun(*sl))
        if isinstance(outs[0], tuple):
            return tuple(numpy.stack([o[j] for o in outs], axis=out_axes) for j in range(len(outs[0])))
        return numpy.stack(outs, axis=out_axes)
    return mapped


class custom_vjp:
    def __init__(self, fun, *a, **k):
        self.fun = fun

    def defvjp(self, fwd, bwd, **k):
        pass

    def __call__(self, *a, **k):
        return self.fun(*a, **k)


def default_backend():
    return "cpu"


def device_count(*a):
    return 1


def local_device_count(*a):
    return 1


def process_index(*a):
    return 0


def process_count(*a):
    return 1


class _Dev:
    id = 0
    platform = "cpu"
    device_kind = "numpy-standin"

    def __repr__(self):
        return "NumpyStandInDevice(0)"


def devices(*a):
    return [_Dev()]


def local_devices(*a):
    return [_Dev()]


def device_put(x, *a, **k):
    return numpy.asarray(x)


def device_get(x):
    return _np.asarray(x)


def block_until_ready(x):
    return x
