"""jax.lax stand-in (single block: collectives are unreachable on the oracle path)."""


def stop_gradient(x):
    return x


def while_loop(cond, body, init):
    v = init
    while cond(v):
        v = body(v)
    return v


def fori_loop(lo, hi, body, init):
    v = init
    for i in range(int(lo), int(hi)):
        v = body(i, v)
    return v


def cond(pred, t, f, *ops):
    return t(*ops) if pred else f(*ops)


def scan(f, init, xs=None, length=None):
    raise NotImplementedError("scan not on the oracle path")


def _collective(*a, **k):
    raise RuntimeError("collective called in single-block oracle")


psum = pmax = pmin = pmean = ppermute = all_gather = axis_index = _collective
