"""Pin oracle/port.py against the reference executed on the stand-in (build container only)."""
import os, sys, time
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import port
from oracle.refharness import run_reference as rr


def setup_from_case(case, num):
    d = case["domain"]
    g = num["conservatives"]["convective_fluxes"]["godunov"]
    return port.Setup(
        cells=tuple(d[a]["cells"] for a in "xyz"),
        domain=tuple(tuple(d[a]["range"]) for a in "xyz"),
        bc={f: case["boundary_conditions"][f]["type"] for f in port.FACES},
        gamma=case["material_properties"]["equation_of_state"]["specific_heat_ratio"],
        nh=num["conservatives"]["halo_cells"],
        recon=g.get("reconstruction_variable", "PRIMITIVE"),
        riemann=g.get("riemann_solver", "HLLC"),
        integrator=num["conservatives"]["time_integration"]["integrator"],
        cfl=num["conservatives"]["time_integration"].get("CFL", 0.5),
    )


def check(name, nsteps=3, **kw):
    case, num = rr.customize(*rr.load_case(name), **kw)
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
    print(f"PIN OK  {name:10s} {kw}  ({time.time()-t0:.1f}s)")


if __name__ == "__main__":
    check("sod", cells=(200, None, None))
    check("sod", cells=(100, None, None), recon="PRIMITIVE", riemann="RUSANOV", integrator="RK2")
    check("riemann2d", cells=(32, 32, None))
    check("riemann2d", cells=(24, 32, None), recon="PRIMITIVE")
    check("riemann2d", cells=(24, 24, None), riemann="RUSANOV", integrator="EULER")
    check("tgv", cells=(16, 16, 16), nsteps=2)
    check("tgv", cells=(12, 16, 20), bc="PERIODIC", nsteps=2)
    check("tgv", cells=(16, 16, 16), bc="PERIODIC", recon="PRIMITIVE", riemann="RUSANOV", nsteps=2)
