"""Pin oracle/port.py against the reference executed on the stand-in (build container only)."""
import json, os, sys, time
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import port
from oracle.refharness import run_reference as rr


def setup_from_case(case, num):
    from tests import helpers
    return helpers.setup_from_json(case, num)


def with_dissipation(case, num, mu=None, bulk=0.0, kappa=None, prandtl=None):
    """Switch the viscous / heat flux of a shipped case on (active_physics + transport block)."""
    case, num = json.loads(json.dumps(case)), json.loads(json.dumps(num))
    tr = case["material_properties"].setdefault("transport", {})
    if mu is not None:
        num["active_physics"]["is_viscous_flux"] = True
        tr["dynamic_viscosity"] = {"model": "CUSTOM", "value": float(mu)}
        tr["bulk_viscosity"] = float(bulk)
    if kappa is not None or prandtl is not None:
        num["active_physics"]["is_heat_flux"] = True
        tr.setdefault("dynamic_viscosity", {"model": "CUSTOM", "value": 0.0})
        tr.setdefault("bulk_viscosity", 0.0)
        tr["thermal_conductivity"] = ({"model": "PRANDTL", "prandtl_number": float(prandtl)} if prandtl is not None
                                      else {"model": "CUSTOM", "value": float(kappa)})
    num["conservatives"].setdefault("dissipative_fluxes", {"reconstruction_stencil": "CENTRAL4",
                                                          "derivative_stencil_center": "CENTRAL4",
                                                          "derivative_stencil_face": "CENTRAL4"})
    return case, num


def check(name, nsteps=3, dissipation=None, **kw):
    case, num = rr.customize(*rr.load_case(name), **kw)
    if dissipation:
        case, num = with_dissipation(case, num, **dissipation)
    t0 = time.time()
    ref = rr.ReferenceRun(case, num)
    s = setup_from_case(case, num)
    sl = (slice(None),) + s.interior
    prims, cons = port.initialize(ref.primitives[sl], s)
    assert np.array_equal(prims, ref.primitives), "init prims"
    assert np.array_equal(cons, ref.conservatives), "init cons"
    dt = port.time_step_size(prims, s)
    assert dt == ref.dt, (dt, ref.dt)
    worst = 0.0
    for n in range(nsteps):
        rec = ref.step()
        mine = {"rhs": [], "prims": [], "cons": []}
        prims, cons, dt = port.step(prims, cons, dt, s, mine)
        for k in ("rhs", "prims", "cons"):
            for st, (a, b) in enumerate(zip(mine[k], rec[k])):
                if not np.array_equal(a, b, equal_nan=True):
                    m = np.isfinite(a) & np.isfinite(b)
                    raise AssertionError(f"{name} {kw} step {n} stage {st} {k}: max diff {np.abs(a[m]-b[m]).max()}")
        assert dt == rec["dt_next"], (dt, rec["dt_next"])
        mr, mp = port.positivity_info(prims, s)
        assert mr == rec["min_density"] and mp == rec["min_pressure"]
        assert np.array_equal(port.totals(cons, s), rec["totals"])
    print(f"PIN OK  {name:10s} {kw} {dissipation or ''}  ({time.time()-t0:.1f}s)")


if __name__ == "__main__":
    check("sod", cells=(200, None, None))
    check("sod", cells=(100, None, None), recon="PRIMITIVE", riemann="RUSANOV", integrator="RK2")
    check("riemann2d", cells=(32, 32, None))
    check("riemann2d", cells=(24, 32, None), recon="PRIMITIVE")
    check("riemann2d", cells=(24, 24, None), riemann="RUSANOV", integrator="EULER")
    check("tgv", cells=(16, 16, 16), nsteps=2)
    check("tgv", cells=(12, 16, 20), bc="PERIODIC", nsteps=2)
    check("tgv", cells=(16, 16, 16), bc="PERIODIC", recon="PRIMITIVE", riemann="RUSANOV", nsteps=2)
    check("tgv", cells=(12, 12, 12), nsteps=2, dissipation=dict(mu=1 / 160, prandtl=0.71))
    check("tgv", cells=(10, 12, 14), bc="PERIODIC", nsteps=2, dissipation=dict(mu=1 / 100, bulk=0.002, kappa=0.05))
    check("riemann2d", cells=(20, 24, None), nsteps=2, dissipation=dict(mu=1e-3))
    check("sod", cells=(80, None, None), nsteps=2, dissipation=dict(mu=2e-3, prandtl=0.7))
