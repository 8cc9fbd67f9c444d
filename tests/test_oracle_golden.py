"""CPU: the NumPy oracle (oracle/port.py) against the fixtures the reference produced."""
import numpy as np
import pytest

from oracle import port
from tests import helpers as H


@pytest.mark.parametrize("name", H.golden_names() + H.generic_golden_names() + H.api_golden_names())
def test_port_reproduces_reference_fixture(name):
    g, case, num = H.load_golden(name)
    s = H.setup_from_json(case, num)
    prims, cons = port.initialize(g["prims0"], s)
    assert np.array_equal(prims, g["prims0_halo"])
    assert np.array_equal(cons, g["cons0_halo"])
    dt = port.time_step_size(prims, s)
    assert dt == float(g["dt0"])
    for a in s.active:
        assert np.array_equal(port.rhs_axis(prims, a, s, cons, dt), g[f"rhs_axis{a}"])
    nsteps = len(g["dt"])
    for n in range(1, nsteps + 1):
        rec = {"rhs": [], "prims": [], "cons": []} if n == 1 else None
        prims, cons, dt = port.step(prims, cons, dt, s, rec)
        if n == 1:
            for k in range(port.RK[s.integrator]["stages"]):
                assert np.array_equal(rec["rhs"][k], g[f"rhs_s{k}"])
                assert np.array_equal(rec["prims"][k], g[f"prims_s{k}"], equal_nan=True)
                assert np.array_equal(rec["cons"][k], g[f"cons_s{k}"], equal_nan=True)
        assert dt == g["dt"][n - 1]
        assert np.array_equal(port.totals(cons, s), g["totals"][n - 1])
        mr, mp = port.positivity_info(prims, s)
        assert mr == g["min_density"][n - 1] and mp == g["min_pressure"][n - 1]
        if f"prims_n{n}" in g:
            assert np.array_equal(prims, g[f"prims_n{n}"], equal_nan=True)
            assert np.array_equal(cons, g[f"cons_n{n}"], equal_nan=True)


def test_sod_physics_against_exact_solution():
    """Independent physics check of the oracle: Sod at t~0.2 vs the exact Riemann solution
    (Toro, ch. 4; same role as src/jaxfluids_thirdparty/exact_riemann_solver in the reference)."""
    s = H.make_setup((200, 1, 1), bc="ZEROGRADIENT")
    x = s.cell_centers()[0]
    rho = np.where(x <= 0.5, 1.0, 0.125)
    p = np.where(x <= 0.5, 1.0, 0.1)
    pr = np.zeros((5, 200, 1, 1))
    pr[0, :, 0, 0], pr[4, :, 0, 0] = rho, p
    prims, cons = port.initialize(pr, s)
    dt, t = port.time_step_size(prims, s), 0.0
    while t < 0.2:
        dt = min(dt, 0.2 - t)
        prims, cons, dtn = port.step(prims, cons, dt, s)
        t += dt
        dt = dtn
    rho_num = prims[0, s.nh:-s.nh, 0, 0]
    # exact star region for the classic Sod problem (gamma = 1.4)
    p_star, u_star = 0.30313017805064707, 0.92745262004895057
    rho_star_L, rho_star_R = 0.42631942817849544, 0.26557371170530725
    x_contact = 0.5 + u_star * 0.2
    mid_L = (x > 0.5 + 0.0 * 0.2 + 0.02) & (x < x_contact - 0.03)
    mid_R = (x > x_contact + 0.03) & (x < 0.5 + 1.7521557320301779 * 0.2 - 0.03)
    assert np.max(np.abs(rho_num[mid_L] - rho_star_L)) < 5e-3
    assert np.max(np.abs(rho_num[mid_R] - rho_star_R)) < 5e-3


@pytest.mark.parametrize("tag", ["lv0", "lv1"])
def test_interpolation_limiter_fixture(tag):
    """positivity/is_interpolation_limiter on a state where it fires thousands of times (fixture generated from
    the reference, oracle/refharness/make_goldens.py:make_limiter_fixture)."""
    import json, copy, os
    g = np.load(os.path.join(H.GOLDEN, "special", "limiter_riemann2d_20x24.npz"))
    case, num = json.loads(str(g[f"case_json_{tag}"])), json.loads(str(g[f"num_json_{tag}"]))
    s = H.setup_from_json(case, num)
    assert s.is_interpolation_limiter and s.limit_velocity == (tag == "lv1")
    with np.errstate(all="ignore"):
        prims, cons = port.initialize(g["user"], s, from_user_buffer=True)
        assert np.array_equal(prims, g[f"prims_halo_{tag}"])
        s0 = copy.copy(s)
        s0.is_interpolation_limiter = False
        changed = sum(int((a != b).sum()) for ax in s.active
                      for a, b in zip(port.reconstruct(prims, ax, s0)[:2], port.reconstruct(prims, ax, s)[:2]))
        assert changed > 1000                                  # the limiter really acts on this state
        assert np.array_equal(port.compute_rhs(prims, s), g[f"rhs_{tag}"], equal_nan=True)


@pytest.mark.parametrize("tag", ["simple", "nasa", "simple_cellsize", "nasa_interp"])
def test_flux_limiter_fixture(tag):
    """positivity/flux_limiter on a state where hundreds of faces fall back to the first-order flux (fixture from the
    reference, oracle/refharness/make_goldens.py:make_flux_limiter_fixture)."""
    import copy, json, os
    g = np.load(os.path.join(H.GOLDEN, "special", "flux_limiter_riemann2d_20x24.npz"))
    s = H.setup_from_json(json.loads(str(g[f"case_json_{tag}"])), json.loads(str(g[f"num_json_{tag}"])))
    prims, cons, dt = g[f"prims_halo_{tag}"], g[f"cons_halo_{tag}"], float(g[f"dt_{tag}"])
    with np.errstate(all="ignore"):
        assert np.array_equal(port.compute_rhs(prims, s, cons, dt), g[f"rhs_{tag}"])
        s0 = copy.copy(s)
        s0.flux_limiter = None
        switched = sum(int((port.face_flux(prims, a, s, cons, dt) != port.face_flux(prims, a, s0)).any(axis=0).sum())
                       for a in s.active)
    assert switched > 80                                   # the limiter really acts on this state


@pytest.mark.parametrize("variable", ["prim", "char"])
@pytest.mark.parametrize("stencil", H.GENERIC_STENCILS)
def test_generic_stencil_fixture(stencil, variable):
    """Every generic reconstruction stencil (WENO1, WENO3-JS/-Z, TENO5, WENO6-CU, the six MUSCL limiters) x
    {PRIMITIVE, CHAR-PRIMITIVE} on one shocked 2-D state: rhs bit-identical to what the reference produced
    (oracle/refharness/make_goldens.py:make_stencil_fixture), and the stencil really differs from WENO5-Z there."""
    import copy, json, os
    g = np.load(os.path.join(H.GOLDEN, "special", "stencils_riemann2d_20x24.npz"))
    key = f"{stencil}_{variable}"
    s = H.setup_from_json(json.loads(str(g[f"case_json_{key}"])), json.loads(str(g[f"num_json_{key}"])))
    assert s.stencil == stencil and s.recon == {"prim": "PRIMITIVE", "char": "CHAR-PRIMITIVE"}[variable]
    prims, cons = port.initialize(g["user"], s, from_user_buffer=True)
    assert np.array_equal(prims, g["prims_halo"])
    rhs = port.compute_rhs(prims, s)
    assert np.array_equal(rhs, g[f"rhs_{key}"])
    s5 = copy.copy(s)
    s5.stencil = "WENO5-Z"
    assert H.rel_linf(port.compute_rhs(prims, s5), rhs) > 1e-3
