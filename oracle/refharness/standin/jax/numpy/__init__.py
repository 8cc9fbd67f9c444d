"""jax.numpy stand-in: NumPy with an immutable-flavoured ndarray subclass."""
import numpy as _np
from numpy import *  # noqa: F401,F403
from numpy import (abs, all, any, max, min, sum, round, bool_, float64, float32, int32, int64,  # noqa: F401
                   uint8, uint32, complex128, newaxis, pi, inf, nan, e, s_, index_exp, linalg, fft)

float_ = _np.float64


class _AtIndexer:
    __slots__ = ("a", "idx")

    def __init__(self, a, idx):
        self.a, self.idx = a, idx

    def _apply(self, fn):
        out = _np.array(self.a, copy=True)
        fn(out)
        return out.view(ndarray)

    def set(self, v, **k):
        def f(o): o[self.idx] = v
        return self._apply(f)

    def add(self, v, **k):
        def f(o): _np.add.at(o, self.idx, v)      # scatter-add: duplicate indices accumulate, like jax
        return self._apply(f)

    def subtract(self, v, **k):
        def f(o): o[self.idx] = o[self.idx] - v
        return self._apply(f)

    def mul(self, v, **k):
        def f(o): o[self.idx] = o[self.idx] * v
        return self._apply(f)

    multiply = mul

    def divide(self, v, **k):
        def f(o): o[self.idx] = o[self.idx] / v
        return self._apply(f)

    def min(self, v, **k):
        def f(o): o[self.idx] = _np.minimum(o[self.idx], v)
        return self._apply(f)

    def max(self, v, **k):
        def f(o): o[self.idx] = _np.maximum(o[self.idx], v)
        return self._apply(f)

    def get(self, **k):
        return self.a[self.idx]


class _At:
    __slots__ = ("a",)

    def __init__(self, a):
        self.a = a

    def __getitem__(self, idx):
        return _AtIndexer(self.a, idx)


class ndarray(_np.ndarray):
    """ndarray whose augmented assignments do not alias (JAX immutability)."""

    @property
    def at(self):
        return _At(self)

    def block_until_ready(self):
        return self

    def __iadd__(self, o): return _np.add(self, o)
    def __isub__(self, o): return _np.subtract(self, o)
    def __imul__(self, o): return _np.multiply(self, o)
    def __itruediv__(self, o): return _np.true_divide(self, o)
    def __ipow__(self, o): return _np.power(self, o)

    def __setitem__(self, k, v):
        raise TypeError("stand-in jax arrays are immutable; use .at[].set()")

    # `array != None` / `array == None` are plain identity tests for a jax.Array (the reference uses them,
    # e.g. halos/outer/material.py:305); NumPy would broadcast them elementwise
    def __ne__(self, o):
        return True if o is None else _np.ndarray.__ne__(self, o)

    def __eq__(self, o):
        return False if o is None else _np.ndarray.__eq__(self, o)

    __hash__ = None


def _wrap(x):
    if isinstance(x, _np.ndarray) and not isinstance(x, ndarray):
        return x.view(ndarray)
    if isinstance(x, tuple):
        return tuple(_wrap(i) for i in x)
    if isinstance(x, list):
        return [_wrap(i) for i in x]
    return x


def _wrapping(fn):
    def w(*a, **k):
        return _wrap(fn(*a, **k))
    w.__name__ = getattr(fn, "__name__", "fn")
    return w


def _red(fn):
    def w(a, axis=None, *args, where=None, initial=None, keepdims=False, **k):
        kw = dict(axis=axis, keepdims=keepdims)
        if where is not None:
            kw["where"] = where
        if initial is not None:
            kw["initial"] = initial
        return _wrap(fn(a, *args, **kw, **k))
    return w


for _n in ("array", "asarray", "zeros", "ones", "empty", "full", "zeros_like", "ones_like", "full_like",
           "arange", "linspace", "stack", "concatenate", "where", "meshgrid", "expand_dims", "squeeze",
           "reshape", "transpose", "swapaxes", "moveaxis", "roll", "flip", "tile", "repeat", "eye",
           "einsum", "matmul", "dot", "cumsum", "clip", "pad", "take", "broadcast_to", "diff", "sort",
           "argsort", "outer", "cross", "tensordot", "identity", "diag", "tril", "triu", "atleast_1d",
           "split", "array_split", "hstack", "vstack", "copy", "real", "imag", "conj", "mean", "prod",
           "argmax", "argmin", "ravel", "rollaxis", "sign", "sqrt", "square", "minimum", "maximum",
           "exp", "log", "sin", "cos", "tan", "tanh", "arctan2", "arctan", "power", "absolute",
           "floor", "ceil", "mod", "logical_and", "logical_or", "logical_not", "isnan", "isinf",
           "count_nonzero", "nonzero", "unique", "interp", "gradient", "trapezoid", "heaviside"):
    if hasattr(_np, _n):
        globals()[_n] = _wrapping(getattr(_np, _n))

abs = _wrapping(_np.abs)
round = _wrapping(_np.round)
min = _red(_np.min)
max = _red(_np.max)
amin = min
amax = max
sum = _red(_np.sum)
all = _red(_np.all)
any = _red(_np.any)


class _FFT:
    """jnp.fft: numpy.fft returning the immutable-flavoured subclass (`buffer_hat /= N**3` inside a callee must not
    alias the caller's array: turbulence/statistics/utilities/energy_spectrum.py:86)."""

    def __getattr__(self, name):
        return _wrapping(getattr(_np.fft, name))


fft = _FFT()


def array(x, dtype=None, **k):
    # jnp.array(list_of_arrays) -> stacked array
    return _wrap(_np.array(x, dtype=dtype))


def asarray(x, dtype=None, **k):
    return _wrap(_np.asarray(x, dtype=dtype))
