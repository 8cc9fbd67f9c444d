"""Per-stage RHS error of the GPU kernels in BOTH norms, on the reference's fixtures: printed as one JSON line.

    python -m tests.tools.strict_norm_check [fixture ...]

  strict  = SURVEY 8(c) / north-star norm: max|a_v - b_v| / max|b_v| per field (fields that are ~0 floored at 1e-3 of
            the largest field) -- `helpers.rel_linf` with its default scales;
  terms   = the same error over the size of the axis contributions the RHS is the sum of (`helpers.rhs_scales`).
Run with JXF_LIB_VARIANT=reforder (a -DJXF_REFERENCE_ORDER -fmad=false build: the reference's operations in the
reference's order) the strict norm must be <= 1e-12: tests/test_gpu_production.py asserts it in a subprocess, which
proves that the only deviation of the production build from the reference is rounding (re-association + FMA).
"""
import json
import sys

import numpy as np
import torch

from oracle import port
from tests import helpers as H
from tests.test_gpu_parity import make_solver, dev, host

DEFAULT = ("tgv16_sym_char_hllc_rk3", "tgv_12x16x20_per_char_hllc_rk3", "riemann2d_32x32_char_hllc_rk3",
           "riemann2d_24x40_prim_hllc_euler", "sod200_char_hllc_rk3", "tgv16_per_char_rusanov_rk3")


def check(name):
    g, case, num = H.load_golden(name)
    s = H.setup_from_json(case, num)
    sol = make_solver(s)
    out = {"strict": 0.0, "terms": 0.0, "axis_strict": 0.0, "step_prims": None}
    nst = port.RK[s.integrator]["stages"]
    for k in range(nst):
        p_in = g["prims0_halo"] if k == 0 else g[f"prims_s{k-1}"]
        p_in = np.nan_to_num(p_in, nan=1.0, posinf=1.0, neginf=1.0)
        got = host(sol.compute_rhs(dev(p_in)))
        ref = g[f"rhs_s{k}"]
        out["strict"] = max(out["strict"], H.rel_linf(got, ref))
        out["terms"] = max(out["terms"], H.rel_linf(got, ref, scale=H.rhs_scales(p_in, s)))
    p0 = dev(g["prims0_halo"])
    tot = g["rhs_s0"]
    for a in s.active:          # one axis' contribution, normalised by the stage rhs it is a term of
        rhs = sol.new_rhs()
        sol.sweep(a, p0, rhs, accumulate=False)
        out["axis_strict"] = max(out["axis_strict"], H.rel_linf(host(rhs), g[f"rhs_axis{a}"], scale=H.field_scales(tot)))
    from jaxfluids_b200.engine import BlockState
    st = BlockState(sol, np.nan_to_num(g["prims0_halo"], nan=1.0, posinf=1.0, neginf=1.0),
                    np.nan_to_num(g["cons0_halo"], nan=1.0, posinf=1.0, neginf=1.0))
    st.step()
    mask = H.face_halo_mask(s)
    out["step_prims"] = H.rel_linf(host(st.primitives)[:, mask], g[f"prims_s{nst-1}"][:, mask])
    out["dt"] = abs(st.dt.item() - g["dt"][0]) / g["dt"][0]
    return out


def main(names):
    res = {n: check(n) for n in (names or DEFAULT)}
    print("STRICT_NORM " + json.dumps(res))


if __name__ == "__main__":
    main(sys.argv[1:])
