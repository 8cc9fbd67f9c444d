"""Generate tests/golden/*.npz from the UNMODIFIED reference sources (build container only).

    python oracle/refharness/make_goldens.py

Each fixture holds what the reference produced for one small case:
  case_json / num_json : the exact setup dicts (so tests rebuild the same setup)
  prims0               : initial primitives, interior (5,Nx,Ny,Nz)
  dt0                  : initial time step
  rhs_axis{a}          : per-axis rhs of the initial state (SpaceSolver.compute_rhs_xi)
  rhs_s{k}             : total rhs at RK stage k of step 0
  prims_s{k}, cons_s{k}: halo'd state after RK stage k of step 0
  prims_n{N}, cons_n{N}: halo'd state after N steps
  dt, time, totals, min_density, min_pressure : per-step sequences
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.refharness import run_reference as rr  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

FIXTURES = {
    # name: (case, kwargs, nsteps, snapshots)
    "sod200_char_hllc_rk3": ("sod", dict(cells=(200, None, None)), 100, (1, 10, 100)),
    "sod100_prim_rusanov_rk2": ("sod", dict(cells=(100, None, None), recon="PRIMITIVE", riemann="RUSANOV",
                                             integrator="RK2"), 10, (1, 10)),
    "riemann2d_32x32_char_hllc_rk3": ("riemann2d", dict(cells=(32, 32, None)), 10, (1, 10)),
    "riemann2d_24x40_prim_hllc_euler": ("riemann2d", dict(cells=(24, 40, None), recon="PRIMITIVE",
                                                           integrator="EULER"), 5, (1, 5)),
    "tgv16_sym_char_hllc_rk3": ("tgv", dict(cells=(16, 16, 16)), 5, (1, 5)),
    "tgv_12x16x20_per_char_hllc_rk3": ("tgv", dict(cells=(12, 16, 20), bc="PERIODIC"), 3, (1, 3)),
    "tgv16_per_char_rusanov_rk3": ("tgv", dict(cells=(16, 16, 16), bc="PERIODIC", riemann="RUSANOV"), 2, (2,)),
    # viscous + heat flux (CENTRAL4 dissipative stencils, edge halos): kwargs carry a `dissipation` entry
    "tgv12_sym_visc_prandtl_rk3": ("tgv", dict(cells=(12, 12, 12), dissipation=dict(mu=1 / 160, prandtl=0.71)), 3, (1, 3)),
    "tgv_10x12x14_per_visc_bulk_kappa_rk3": ("tgv", dict(cells=(10, 12, 14), bc="PERIODIC",
                                                         dissipation=dict(mu=1 / 100, bulk=0.002, kappa=0.05)), 2, (2,)),
    "riemann2d_20x24_visc_rk3": ("riemann2d", dict(cells=(20, 24, None), dissipation=dict(mu=1e-3)), 3, (3,)),
    "sod80_visc_prandtl_rk3": ("sod", dict(cells=(80, None, None), dissipation=dict(mu=2e-3, prandtl=0.7)), 5, (5,)),
    # WENO5-JS
    "sod100_js_char_hllc_rk3": ("sod", dict(cells=(100, None, None), stencil="WENO5-JS"), 20, (1, 20)),
    "riemann2d_24x28_js_prim_hllc_rk3": ("riemann2d", dict(cells=(24, 28, None), stencil="WENO5-JS", recon="PRIMITIVE"), 5, (5,)),
    "tgv_10x12x10_per_js_prim_visc_rk3": ("tgv", dict(cells=(10, 12, 10), bc="PERIODIC", stencil="WENO5-JS", recon="PRIMITIVE",
                                                      dissipation=dict(mu=1e-2)), 2, (2,)),
    # HLL Riemann solver (solvers/riemann_solvers/HLL.py)
    "sod100_char_hll_rk3": ("sod", dict(cells=(100, None, None), riemann="HLL"), 10, (10,)),
    "riemann2d_20x24_prim_hll_davis_rk3": ("riemann2d", dict(cells=(20, 24, None), recon="PRIMITIVE", riemann="HLL",
                                                               signal_speed="DAVIS"), 3, (3,)),
    # the shipped double-rarefaction example, shrunk: flux limiter SIMPLE + interpolation limiter, WENO5-JS, nh 4
    "rarefaction100_fluxlim_simple_rk3": ("rarefaction", dict(cells=(100, None, None)), 30, (30,)),
    # ... and a stronger one (Mach 9.4 apart): ~500 face fluxes replaced over the 40 steps
    "rarefaction100_strong_fluxlim_simple_rk3": ("rarefaction", dict(cells=(100, None, None), initial_condition={
        "u": "lambda x: -2.5*(x <= 0.5) + 2.5*(x > 0.5)", "p": 0.05}), 40, (40,)),
    "riemann2d_16x20_fluxlim_nasa_cellsize_rk3": ("riemann2d", dict(cells=(16, 20, None), positivity={
        "flux_limiter": "NASA", "flux_partition": "CELLSIZE"}), 3, (3,)),
    # the shipped lid-driven cavity example, shrunk: WALL on four faces (moving lid), WENO5-JS PRIMITIVE, viscous,
    # interpolation limiter on, halo_cells 4
    "cavity_24x20_wall_js_visc_rk3": ("cavity", dict(cells=(24, 20, None)), 5, (5,)),
    # the shipped Rayleigh-Taylor example, shrunk: DIRICHLET north/south, SYMMETRY east/west, gravity, limiter
    "rti_16x48_dirichlet_gravity_rk3": ("rti", dict(cells=(16, 48, None)), 5, (5,)),
    # the shipped 1-D heat equation example, shrunk: heat flux only (is_convective_flux false), DIRICHLET E/W, nh 4
    "heat1d_40_dirichlet_noconv_rk3": ("heat1d", dict(cells=(40, None, None)), 6, (6,)),
    # RK2_LS4 (time_integration/RK2_LS4.py) and the generic reconstruction stencils (TENO5, WENO3-Z, WENO6-CU, MUSCL);
    # kept in tests/golden/generic/ (tests/helpers.generic_golden_names)
    "generic/sod100_teno5_char_hllc_rk2ls4": ("sod", dict(cells=(100, None, None), stencil="TENO5", integrator="RK2_LS4"), 20, (20,)),
    "generic/riemann2d_20x24_weno3z_prim_hllc_rk3": ("riemann2d", dict(cells=(20, 24, None), stencil="WENO3-Z", recon="PRIMITIVE"), 3, (3,)),
    "generic/riemann2d_16x20_weno6cu_char_hllc_rk3": ("riemann2d", dict(cells=(16, 20, None), stencil="WENO6-CU"), 3, (3,)),
    "generic/sod100_vanleer_prim_rusanov_rk3": ("sod", dict(cells=(100, None, None), stencil="VANLEER", recon="PRIMITIVE",
                                                     riemann="RUSANOV"), 10, (10,)),
    "generic/tgv_10x8x12_per_minmod_char_hllc_rk2ls4": ("tgv", dict(cells=(10, 8, 12), bc="PERIODIC", stencil="MINMOD",
                                                            integrator="RK2_LS4"), 2, (2,)),
    "generic/riemann2d_16x20_teno6_char_hllc_rk3": ("riemann2d", dict(cells=(16, 20, None), stencil="TENO6"), 3, (3,)),
    # convective_solver = FLUX-SPLITTING (flux_splitting_scheme.py): the shipped Lax and Woodward-Colella examples,
    # shrunk, and 2-D / 3-D variants with the other eigenvalue choices
    "generic/lax100_fs_roe_weno6cu_rk3": ("lax", dict(cells=(100, None, None)), 20, (20,)),
    "generic/woodward200_fs_roe_weno5z_rk3": ("woodward", dict(cells=(200, None, None)), 30, (30,)),
    "generic/riemann2d_16x20_fs_cllf_teno5_rk3": ("riemann2d", dict(cells=(16, 20, None), flux_splitting="CLLF",
                                                                    stencil="TENO5"), 3, (3,)),
    "generic/tgv_10x8x12_fs_llf_weno5js_rk3": ("tgv", dict(cells=(10, 8, 12), flux_splitting="LLF", stencil="WENO5-JS"), 2, (2,)),
    # reconstruction_variable CONSERVATIVE / CHAR-CONSERVATIVE and frozen_state ROE
    "generic/sod100_charcons_roe_hllc_rk3": ("sod", dict(cells=(100, None, None), recon="CHAR-CONSERVATIVE", frozen_state="ROE"), 10, (10,)),
    "generic/riemann2d_16x20_cons_teno5_hll_rk3": ("riemann2d", dict(cells=(16, 20, None), recon="CONSERVATIVE", stencil="TENO5",
                                                                     riemann="HLL"), 3, (3,)),
    "generic/tgv_10x8x12_char_roe_hllc_rk3": ("tgv", dict(cells=(10, 8, 12), frozen_state="ROE"), 2, (2,)),
    "generic/tgv_8x10x12_per_charcons_hllc_rk3": ("tgv", dict(cells=(8, 10, 12), bc="PERIODIC", recon="CHAR-CONSERVATIVE"), 2, (2,)),
    "generic/lax100_fs_roe_weno6cu_roefrozen_rk3": ("lax", dict(cells=(100, None, None), frozen_state="ROE"), 10, (10,)),
    # the shipped 2-D heat equation example, shrunk: heat flux only, DIRICHLET on four faces with p(x) = 1 + sin(pi x) at
    # the north face (a lambda in the case file).  tests/golden/api/: fixtures that run through the public API only
    "api/heat2d_24x20_dirichlet_lambda_noconv_rk3": ("heat2d", dict(cells=(24, 20, None)), 6, (6,)),
    # SIMPLE_INFLOW (west) / SIMPLE_OUTFLOW (east) / NEUMANN (north, south) with constants and lambdas, convective and
    # with the viscous + heat flux (edge halos)
    "api/riemann2d_16x20_inflow_outflow_neumann_rk3": ("riemann2d", dict(cells=(16, 20, None), boundary_conditions={
        "west": {"type": "SIMPLE_INFLOW", "primitives_callable": {"rho": "lambda y,t: 0.5 + 0.2 * y", "u": 1.2,
                                                                  "v": "lambda y,t: 0.1 * jnp.sin(6 * y)", "w": 0.0}},
        "east": {"type": "SIMPLE_OUTFLOW", "primitives_callable": {"p": "lambda y,t: 1.0 + 0.5 * y"}},
        "north": {"type": "NEUMANN", "primitives_callable": {"rho": 0.1, "u": "lambda x,t: 0.2 * x", "v": 0.0, "w": 0.0,
                                                             "p": -0.3}},
        "south": {"type": "NEUMANN", "primitives_callable": {"rho": "lambda x,t: 0.3 * jnp.cos(5 * x)", "u": 0.0, "v": 0.1,
                                                             "w": 0.0, "p": 0.2}}}), 4, (4,)),
    "api/riemann2d_16x20_inflow_outflow_visc_rk3": ("riemann2d", dict(cells=(16, 20, None), dissipation=dict(mu=1e-3, kappa=1e-3),
                                                                      boundary_conditions={
        "west": {"type": "SIMPLE_INFLOW", "primitives_callable": {"rho": 1.0, "u": 0.3, "v": 0.0, "w": 0.0}},
        "east": {"type": "SIMPLE_OUTFLOW", "primitives_callable": {"p": 0.1}},
        "north": {"type": "NEUMANN", "primitives_callable": {"rho": 0.1, "u": "lambda x,t: 0.2 * x", "v": 0.0, "w": 0.0,
                                                             "p": -0.3}}}), 3, (3,)),
    # the shipped double Mach reflection example, shrunk: the south face is DIRICHLET for x < 1/6 and SYMMETRY beyond
    "api/dmr_48x32_dirichlet_symmetry_south_rk3": ("dmr", dict(cells=(48, 32, None)), 8, (8,)),
    "generic/sod100_teno6a_char_hllc_rk3": ("sod", dict(cells=(100, None, None), stencil="TENO6-A"), 10, (10,)),
    "generic/riemann2d_16x20_teno5a_prim_hllc_rk3": ("riemann2d", dict(cells=(16, 20, None), stencil="TENO5-A", recon="PRIMITIVE"), 3, (3,)),
    # the shipped lid-driven cavity with a regularised lid u(x) = 16 x^2 (1 - x)^2: WALL with a space-dependent velocity
    "api/cavity_24x20_wall_lambda_lid_visc_rk3": ("cavity", dict(cells=(24, 20, None), boundary_conditions={
        "north": {"type": "WALL", "wall_velocity_callable": {"u": "lambda x,t: 16.0 * x**2 * (1.0 - x)**2", "v": 0.0,
                                                             "w": 0.0}}}), 4, (4,)),
    # HLLC-LM (HLLCLM.py; the TGV at Mach 0.1 is where its low-Mach limiter acts) and AUSM+ (AUSMP.py)
    "generic/sod100_char_hllclm_rk3": ("sod", dict(cells=(100, None, None), riemann="HLLC-LM"), 10, (10,)),
    "generic/tgv_10x8x12_sym_char_hllclm_rk3": ("tgv", dict(cells=(10, 8, 12), riemann="HLLC-LM"), 2, (2,)),
    "generic/riemann2d_20x24_prim_ausmp_rk3": ("riemann2d", dict(cells=(20, 24, None), recon="PRIMITIVE", riemann="AUSMP"), 3, (3,)),
    "generic/sod100_char_ausmp_weno3z_rk2": ("sod", dict(cells=(100, None, None), riemann="AUSMP", stencil="WENO3-Z",
                                                          integrator="RK2"), 10, (10,)),
}

GENERIC_STENCILS = ("WENO1", "WENO3-JS", "WENO3-Z", "WENO3-N", "CENTRAL2", "TENO5", "TENO5-A", "TENO6", "TENO6-A", "WENO6-CU",
                    "KOREN", "MC", "MINMOD", "SUPERBEE", "VANALBADA", "VANLEER")


def make_stencil_fixture():
    """rhs of one shocked, oscillatory 2-D state for every generic reconstruction stencil x {PRIMITIVE, CHAR-PRIMITIVE}:
    jumps (the TENO cut-off and the slope limiters switch) next to smooth waves of both signs."""
    d = {}
    x, y = np.meshgrid(np.linspace(0, 1, 20), np.linspace(0, 1, 24), indexing="ij")
    rho = np.where(x < 0.45, 1.0 + 0.3 * np.sin(9 * y), 0.25 + 0.1 * np.cos(7 * x + 5 * y)) + 0.5 * (y > 0.6)
    p = np.where(x + y < 0.9, 1.0 + 0.2 * np.cos(5 * x), 0.15 + 0.05 * np.sin(11 * y))
    user = np.stack([rho, 0.6 * np.sin(6 * x + 2 * y), -0.4 * np.cos(4 * y - 3 * x) + 0.3 * (x > 0.7), p])[..., None]
    d["user"] = user
    for st in GENERIC_STENCILS:
        for rv, tag in (("PRIMITIVE", "prim"), ("CHAR-PRIMITIVE", "char")):
            case, num = rr.customize(*rr.load_case("riemann2d"), cells=(20, 24, None), stencil=st, recon=rv)
            run = rr.ReferenceRun(case, num, user_prime_init=user)
            key = f"{st}_{tag}"
            d[f"case_json_{key}"], d[f"num_json_{key}"] = np.array(json.dumps(case)), np.array(json.dumps(num))
            d["prims_halo"] = run.primitives
            d[f"rhs_{key}"] = run.compute_rhs()
    path = os.path.join(OUT, "special", "stencils_riemann2d_20x24.npz")
    np.savez_compressed(path, **d)
    print(f"special/stencils_riemann2d_20x24: {os.path.getsize(path) / 1e6:.2f} MB")


def make_limiter_fixture():
    """rhs of a state on which the interpolation limiter (limiter_interpolation.py:77-209) fires thousands of
    times (density ~3e-13 / pressure ~4e-11 regions), for both settings of positivity/limit_velocity."""
    d = {}
    for lv in (False, True):
        case, num = rr.customize(*rr.load_case("riemann2d"), cells=(20, 24, None), recon="CHAR-PRIMITIVE")
        num["conservatives"]["positivity"] = {"is_interpolation_limiter": True, "limit_velocity": lv}
        x, y = np.meshgrid(np.linspace(0, 1, 20), np.linspace(0, 1, 24), indexing="ij")
        rho = np.where(x < 0.5, 1.0 + 0.1 * np.sin(7 * y), 3e-13 * (1 + 0.5 * np.sin(9 * y + 3 * x)))
        p = np.where(y < 0.5, 1.0 + 0.2 * np.cos(5 * x), 4e-11 * (1 + 0.5 * np.cos(11 * x + y)))
        user = np.stack([rho, 0.3 * np.sin(3 * x), 0.2 * np.cos(4 * y), p])[..., None]
        run = rr.ReferenceRun(case, num, user_prime_init=user)
        tag = "lv1" if lv else "lv0"
        d[f"case_json_{tag}"], d[f"num_json_{tag}"] = np.array(json.dumps(case)), np.array(json.dumps(num))
        d["user"] = user
        d[f"prims_halo_{tag}"] = run.primitives
        d[f"rhs_{tag}"] = run.compute_rhs()
    os.makedirs(os.path.join(OUT, "special"), exist_ok=True)
    path = os.path.join(OUT, "special", "limiter_riemann2d_20x24.npz")
    np.savez_compressed(path, **d)
    print(f"special/limiter_riemann2d_20x24: {os.path.getsize(path) / 1e6:.2f} MB")


def make_flux_limiter_fixture():
    """rhs of a near-vacuum, fast-moving state on which the positivity flux limiter (limiter_flux.py:146-330) replaces
    hundreds of face fluxes by the first-order flux -- per variant (SIMPLE / NASA, UNIFORM / CELLSIZE partition)."""
    d = {}
    x, y = np.meshgrid(np.linspace(0, 1, 20), np.linspace(0, 1, 24), indexing="ij")
    rho = np.where(x < 0.5, 1.0 + 0.1 * np.sin(7 * y), 2e-3 * (1 + 0.5 * np.sin(9 * y + 3 * x)))
    p = np.where(y < 0.5, 1.0 + 0.2 * np.cos(5 * x), 1e-3 * (1 + 0.5 * np.cos(11 * x + y)))
    user = np.stack([rho, 3.0 * np.sin(3 * x + 2 * y), 2.0 * np.cos(4 * y - x), p])[..., None]
    d["user"] = user
    for tag, pos in (("simple", {"flux_limiter": "SIMPLE"}), ("nasa", {"flux_limiter": "NASA"}),
                     ("simple_cellsize", {"flux_limiter": "SIMPLE", "flux_partition": "CELLSIZE"}),
                     ("nasa_interp", {"flux_limiter": "NASA", "is_interpolation_limiter": True})):
        case, num = rr.customize(*rr.load_case("riemann2d"), cells=(20, 24, None), positivity=pos)
        case["domain"]["y"]["range"] = [0.0, 1.7]                # dx != dy: the partitions differ
        run = rr.ReferenceRun(case, num, user_prime_init=user)
        d[f"case_json_{tag}"], d[f"num_json_{tag}"] = np.array(json.dumps(case)), np.array(json.dumps(num))
        d[f"prims_halo_{tag}"], d[f"cons_halo_{tag}"] = run.primitives, run.conservatives
        # a time step 6x the CFL one makes the pseudo-integration overshoot on many faces
        d[f"dt_{tag}"] = np.float64(6.0 * run.dt)
        d[f"rhs_{tag}"] = run.compute_rhs(dt=6.0 * run.dt)
    path = os.path.join(OUT, "special", "flux_limiter_riemann2d_20x24.npz")
    np.savez_compressed(path, **d)
    print(f"special/flux_limiter_riemann2d_20x24: {os.path.getsize(path) / 1e6:.2f} MB")


def make_hit_fixture():
    """SURVEY 8(c) / BASELINE config 5 at 32^3: periodic [0, 2 pi]^3, gamma 1.4, the benchmark's synthetic solenoidal
    field (jaxfluids_b200/turbulence.synthetic_solenoidal_ic: E(k) ~ k^4 exp(-2 k^2 / k0^2), k0 = 4, Ma_t = 0.4,
    seed 0) injected through initialization(user_prime_init=...), TGV numerical setup (CHAR-PRIMITIVE WENO5-Z + HLLC +
    RK3).  Compact: interior arrays only (the IC is stored, not regenerated: its BLAS contractions are not
    bit-reproducible across hosts)."""
    from jaxfluids_b200 import turbulence
    case, num = rr.customize(*rr.load_case("tgv"), cells=(32, 32, 32), bc="PERIODIC")
    case["material_properties"]["equation_of_state"]["specific_heat_ratio"] = 1.4
    user = turbulence.synthetic_solenoidal_ic(32, gamma=1.4, k0=4.0, ma_t=0.4, seed=0)
    run = rr.ReferenceRun(case, num, user_prime_init=user)
    d = {"case_json": np.array(json.dumps(case)), "num_json": np.array(json.dumps(num)), "user": user,
         "dt0": np.float64(run.dt)}
    nsteps = 3
    seq = {"dt": [], "time": [], "totals": [], "min_density": [], "min_pressure": []}
    for n in range(1, nsteps + 1):
        rec = run.step(record_stages=(n == 1))
        if n == 1:
            d["rhs_s0"] = rec["rhs"][0]
        for k, key in (("dt", "dt_next"), ("time", "time"), ("totals", "totals"), ("min_density", "min_density"),
                       ("min_pressure", "min_pressure")):
            seq[k].append(rec[key])
    d[f"prims_n{nsteps}"] = run.interior(run.primitives).copy()
    for k, v in seq.items():
        d[k] = np.array(v)
    path = os.path.join(OUT, "special", "hit32_per_char_hllc_rk3.npz")
    np.savez_compressed(path, **d)
    print(f"special/hit32_per_char_hllc_rk3: {os.path.getsize(path) / 1e6:.2f} MB")


def make(name, case_name, kw, nsteps, snaps):
    kw = dict(kw)
    dissipation = kw.pop("dissipation", None)
    case, num = rr.customize(*rr.load_case(case_name), **kw)
    if dissipation:
        from oracle.refharness import pin_check
        case, num = pin_check.with_dissipation(case, num, **dissipation)
    run = rr.ReferenceRun(case, num)
    d = {"case_json": np.array(json.dumps(case)), "num_json": np.array(json.dumps(num))}
    d["prims0"] = run.interior(run.primitives).copy()
    d["prims0_halo"] = run.primitives
    d["cons0_halo"] = run.conservatives
    d["dt0"] = np.float64(run.dt)
    # per-axis rhs of the initial state
    ss = run.sim.space_solver
    mf = run.material_fields
    for a in run.sim.domain_information.active_axes_indices:
        out = ss.compute_rhs_xi(mf.conservatives, mf.primitives, mf.temperature, a, 0.0, run.dt,
                                ml_setup=run.default_ml_setup())
        d[f"rhs_axis{a}"] = np.array(out[0].conservatives)
    seq = {"dt": [], "time": [], "totals": [], "min_density": [], "min_pressure": []}
    for n in range(1, nsteps + 1):
        rec = run.step(record_stages=(n == 1))
        if n == 1:
            for k, (r, p, c) in enumerate(zip(rec["rhs"], rec["prims"], rec["cons"])):
                d[f"rhs_s{k}"], d[f"prims_s{k}"], d[f"cons_s{k}"] = r, p, c
        seq["dt"].append(rec["dt_next"])
        seq["time"].append(rec["time"])
        seq["totals"].append(rec["totals"])
        seq["min_density"].append(rec["min_density"])
        seq["min_pressure"].append(rec["min_pressure"])
        if n in snaps:
            d[f"prims_n{n}"], d[f"cons_n{n}"] = run.primitives, run.conservatives
    for k, v in seq.items():
        d[k] = np.array(v)
    path = os.path.join(OUT, name + ".npz")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.savez_compressed(path, **d)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]
    for name, (case_name, kw, nsteps, snaps) in FIXTURES.items():
        if only and name not in only:
            continue
        with np.errstate(all="ignore"):
            make(name, case_name, kw, nsteps, snaps)
    if not only or "limiter" in only:
        with np.errstate(all="ignore"):
            make_limiter_fixture()
    if not only or "flux_limiter" in only:
        with np.errstate(all="ignore"):
            make_flux_limiter_fixture()
    if not only or "hit" in only:
        with np.errstate(all="ignore"):
            make_hit_fixture()
    if not only or "stencils" in only:
        with np.errstate(all="ignore"):
            make_stencil_fixture()
