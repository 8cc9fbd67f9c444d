"""Slab-threaded driver of the NumPy oracle (oracle/port.py) -- CPU baseline that uses all host
cores.  TEST / BENCH INFRASTRUCTURE ONLY (see the header of port.py).

The grid is cut into slabs along x; every slab evaluates port.compute_rhs on its own view of the
halo'd primitives (a block [a,b) of interior cells is the halo'd range [a, b+2nh)), and the
elementwise stage update / primitive recovery are slab-parallel too.  NumPy releases the GIL inside
ufunc loops, so plain threads scale.  Results are bit-identical to the single-threaded port (same
elementwise operations on the same values), which tests/test_oracle_golden.py asserts.
"""
from __future__ import annotations

import copy
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import port


class ThreadedStepper:
    def __init__(self, s: port.Setup, threads: int | None = None):
        self.s = s
        n0 = s.cells[0]
        assert n0 > 1, "slab threading cuts the x axis"
        self.threads = max(1, min(threads or os.cpu_count() or 1, n0 // 4 or 1))
        edges = np.linspace(0, n0, self.threads + 1).astype(int)
        self.slabs = [(int(a), int(b)) for a, b in zip(edges[:-1], edges[1:]) if b > a]
        self.subs = []
        for a, b in self.slabs:
            sub = copy.copy(s)
            sub.cells = (b - a,) + tuple(s.cells[1:])
            sub.inv_dx_override = tuple(float(v) for v in s.inv_dx)
            sub.active = s.active
            self.subs.append(sub)
        X = s.shape[1]
        hedges = np.linspace(0, X, len(self.slabs) + 1).astype(int)
        self.hslabs = [(int(a), int(b)) for a, b in zip(hedges[:-1], hedges[1:])]
        self.pool = ThreadPoolExecutor(max_workers=len(self.slabs))

    def compute_rhs(self, prims):
        s, nh = self.s, self.s.nh
        out = np.empty((5,) + s.cells)

        def work(i):
            a, b = self.slabs[i]
            out[:, a:b] = port.compute_rhs(prims[:, a:b + 2 * nh], self.subs[i])
        list(self.pool.map(work, range(len(self.slabs))))
        return out

    def stage(self, prims, cons, cons_n, dt, k):
        s = self.s
        rk = port.RK[s.integrator]
        rhs = self.compute_rhs(prims)
        step = dt * rk["dt_mult"][k]
        new_cons = np.empty_like(cons)
        new_prims = np.empty_like(prims)
        nh = s.nh
        inter = s.interior
        X = s.shape[1]

        def work(i):
            a, b = self.hslabs[i]
            c = cons[:, a:b]
            if k > 0:
                ca, cb = rk["blend"][k - 1]
                c = ca * c + cb * cons_n[:, a:b]
            else:
                c = c.copy()
            # interior part of this halo'd x-range
            ia, ib = max(a, nh), min(b, X - nh)
            if ib > ia:
                sl = (slice(None), slice(ia - a, ib - a)) + inter[1:]
                c[sl] = c[sl] + step * rhs[:, ia - nh:ib - nh]
            new_cons[:, a:b] = c
            new_prims[:, a:b] = port.prims_from_cons(c, s.gamma)
        list(self.pool.map(work, range(len(self.hslabs))))
        new_prims, new_cons = port.halo_fill(new_prims, new_cons, s)
        return new_prims, new_cons, rhs

    def step(self, prims, cons, dt):
        cons_n = cons
        with np.errstate(all="ignore"):
            for k in range(port.RK[self.s.integrator]["stages"]):
                prims, cons, _ = self.stage(prims, cons, cons_n, dt, k)
        return prims, cons, port.time_step_size(prims, self.s)
