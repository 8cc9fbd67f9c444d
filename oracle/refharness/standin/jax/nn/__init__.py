import numpy as _np


def relu(x):
    return _np.maximum(x, 0)
