"""GPU parity of the run-time variants that ride on the non-tuned kernel instantiations: the generic reconstruction
stencils (WENO1, WENO3-JS/-Z/-N, CENTRAL2, TENO5, TENO6, WENO6-CU, the MUSCL limiters; STENCIL_GENERIC instantiations),
the HLLC-LM and AUSM+ Riemann solvers (RIEMANN_RUSANOV instantiations) and the RK2_LS4 integrator -- called through the
C ABI, against the pinned CPU oracle and the fixtures the reference produced (tests/golden/generic/,
tests/golden/special/stencils_*.npz).

The device functions behind these tests are host-simulated on the CPU in tests/test_hostsim.py (bit-identical to the
reference without FMA contraction)."""
import json
import os

import numpy as np
import pytest

from oracle import port
from tests import helpers as H
from tests import test_gpu_parity as P

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", H.generic_golden_names())
def test_generic_fixture_rhs_and_stages(name):
    """Per-axis rhs, stage rhs and the first full step of the reference's fixtures (TENO5 + RK2_LS4, WENO3-Z,
    WENO6-CU, TENO6, VANLEER + Rusanov, MINMOD 3-D + RK2_LS4, HLLC-LM 1-D / 3-D low Mach, AUSM+ 2-D, AUSM+ + WENO3-Z)."""
    P.test_reference_fixture_rhs_and_stages(name)


@pytest.mark.parametrize("name", H.generic_golden_names())
def test_generic_fixture_multi_step(name):
    """dt sequence, totals, min rho / min p and the state after N steps against the reference."""
    P.test_reference_fixture_multi_step(name)


@pytest.mark.parametrize("variable", ["prim", "char"])
@pytest.mark.parametrize("stencil", H.GENERIC_STENCILS)
def test_generic_stencil_shock_fixture_rhs(stencil, variable):
    """Every generic stencil x reconstruction variable on the reference's shocked 2-D state (the TENO cut-off and the
    slope limiters switch there): stage rhs within 1e-12 of what the reference produced."""
    g = np.load(os.path.join(H.GOLDEN, "special", "stencils_riemann2d_20x24.npz"))
    key = f"{stencil}_{variable}"
    s = H.setup_from_json(json.loads(str(g[f"case_json_{key}"])), json.loads(str(g[f"num_json_{key}"])))
    sol = P.make_solver(s)
    prims = g["prims_halo"]
    got = P.host(sol.compute_rhs(P.dev(np.nan_to_num(prims, nan=1.0))))
    assert np.isfinite(got).all()
    assert H.rel_linf(got, g[f"rhs_{key}"], scale=H.rhs_scales(prims, s)) <= H.TOL_RHS


@pytest.mark.parametrize("force_rows", [False, True])
@pytest.mark.parametrize("stencil,recon,riemann", [("TENO5", "CHAR-PRIMITIVE", "HLLC"), ("WENO6-CU", "PRIMITIVE", "HLLC"),
                                                   ("WENO3-Z", "CHAR-PRIMITIVE", "RUSANOV"), ("VANALBADA", "PRIMITIVE", "HLL"),
                                                   ("WENO1", "CHAR-PRIMITIVE", "HLLC"), ("TENO6", "CHAR-PRIMITIVE", "HLLC-LM"),
                                                   ("WENO3-N", "PRIMITIVE", "AUSMP")])
def test_generic_stencils_3d_all_kernels(stencil, recon, riemann, force_rows, monkeypatch):
    """The generic instantiations of every sweep kernel (march x / y, rows (TMA) or contig z, fused epilogue): per-axis
    rhs and 3 RK2_LS4 steps against the oracle."""
    from jaxfluids_b200.engine import BlockState
    if force_rows:
        monkeypatch.setenv("JXF_FORCE_ROWS", "1")
    s = H.make_setup((18, 16, 40), bc="SYMMETRY", recon=recon, riemann=riemann, stencil=stencil, integrator="RK2_LS4")
    prims, cons = port.initialize(H.smooth_ic(s, seed=21, amp=0.15), s)
    sol = P.make_solver(s)
    assert sol.stages == 4
    p = P.dev(np.nan_to_num(prims, nan=1.0))
    scales = H.rhs_scales(prims, s)
    for a in s.active:
        rhs = sol.new_rhs()
        sol.sweep(a, p, rhs, accumulate=False)
        assert H.rel_linf(P.host(rhs), port.rhs_axis(prims, a, s), scale=scales) <= H.TOL_RHS, f"axis {a}"
    st = BlockState(sol, np.nan_to_num(prims, nan=1.0), np.nan_to_num(cons, nan=1.0))
    dt = port.time_step_size(prims, s)
    for _ in range(3):
        prims, cons, dt = port.step(prims, cons, dt, s)
        st.step()
    m = H.defined_mask(s)
    assert H.rel_linf(P.host(st.primitives)[:, m], prims[:, m]) <= 1e-12
    assert H.rel_linf(P.host(st.conservatives)[:, m], cons[:, m]) <= 1e-12
    assert abs(st.dt.item() - dt) <= 1e-12 * dt


@pytest.mark.parametrize("cells,bc", [((120, 1, 1), "ZEROGRADIENT"), ((28, 36, 1), "PERIODIC"), ((16, 12, 20), "SYMMETRY")])
def test_rk2_ls4_steps_with_the_tuned_kernels(cells, bc):
    """integrator = RK2_LS4 (RK2_LS4.py: four stages, each restarting from U^n) on the WENO5-Z kernels: 5 steps
    against the oracle, conserved totals to 1e-12."""
    from jaxfluids_b200.engine import BlockState
    s = H.make_setup(cells, bc=bc, integrator="RK2_LS4")
    prims, cons = port.initialize(H.smooth_ic(s, seed=5, amp=0.15), s)
    sol = P.make_solver(s)
    st = BlockState(sol, np.nan_to_num(prims, nan=1.0), np.nan_to_num(cons, nan=1.0))
    dt = port.time_step_size(prims, s)
    for _ in range(5):
        prims, cons, dt = port.step(prims, cons, dt, s)
        st.step()
    m = H.defined_mask(s)
    assert H.rel_linf(P.host(st.primitives)[:, m], prims[:, m]) <= 1e-12
    assert abs(st.dt.item() - dt) <= 1e-12 * dt
    sl = (slice(None),) + s.interior
    tot, ref = P.host(st.conservatives)[sl].reshape(5, -1).sum(axis=1), port.totals(cons, s)
    scale = np.maximum(np.abs(ref), 1e-3 * np.max(np.abs(ref)))
    assert np.max(np.abs(tot - ref) / scale) <= H.TOL_TOTALS * max(1.0, np.sqrt(np.prod(s.cells)) / 10)


def test_public_api_with_generic_stencil_and_rk2_ls4():
    """The reference's JSON with reconstruction_stencil = TENO5 and integrator = RK2_LS4 through InputManager /
    InitializationManager / SimulationManager.simulate, against the fixture the reference produced."""
    from jaxfluids_b200 import InputManager, InitializationManager, SimulationManager
    name = "generic/sod100_teno5_char_hllc_rk2ls4"
    g, case, num = H.load_golden(name)
    n = len(g["dt"])
    case = json.loads(json.dumps(case))
    case["general"]["end_step"] = n
    case["general"]["end_time"] = 1e300
    num = json.loads(json.dumps(num))
    num.setdefault("output", {}).setdefault("logging", {})["level"] = "NONE"
    im = InputManager(case, num)
    buffers = InitializationManager(im).initialization()
    sim = SimulationManager(im)
    assert sim.time_integrator.no_stages == 4
    sim.simulate(buffers)
    out = sim.final_buffers
    s = H.setup_from_json(case, num)
    m = H.face_halo_mask(s)
    pr = P.host(out.simulation_buffers.material_fields.primitives)
    assert out.time_control_variables.simulation_step == n
    assert H.rel_linf(pr[:, m], g[f"prims_n{n}"][:, m]) <= H.TOL_PRIMS_100
    assert abs(out.time_control_variables.physical_simulation_time - g["time"][n - 1]) <= 1e-12 * g["time"][n - 1]


@pytest.mark.parametrize("riemann,sig", [("HLLC-LM", "EINFELDT"), ("HLLC-LM", "TORO"), ("AUSMP", "EINFELDT")])
@pytest.mark.parametrize("cells,bc,recon,factor", [((120, 1, 1), "ZEROGRADIENT", "CHAR-PRIMITIVE", 4.0),
                                                    ((32, 36, 1), "PERIODIC", "PRIMITIVE", 4.0),
                                                    ((16, 12, 40), "SYMMETRY", "CHAR-PRIMITIVE", 0.05)])
def test_hllclm_and_ausmp(cells, bc, recon, factor, riemann, sig):
    """riemann_solver = HLLC-LM (HLLCLM.py) / AUSMP (AUSMP.py) on the tuned WENO5-Z reconstruction: per-axis rhs and 3
    steps against the oracle, on states with supersonic faces of both signs and on a low-Mach state (where the
    HLLC-LM wave-speed limiter acts)."""
    from jaxfluids_b200.engine import BlockState
    s = H.make_setup(cells, bc=bc, recon=recon, riemann=riemann)
    s.signal_speed = sig
    ic = H.smooth_ic(s, seed=4, amp=0.2)
    ic[1:4] *= factor
    prims, cons = port.initialize(ic, s)
    sol = P.make_solver(s)
    p = P.dev(np.nan_to_num(prims, nan=1.0))
    scales = H.rhs_scales(prims, s)
    for a in s.active:
        rhs = sol.new_rhs()
        sol.sweep(a, p, rhs, accumulate=False)
        assert H.rel_linf(P.host(rhs), port.rhs_axis(prims, a, s), scale=scales) <= H.TOL_RHS, f"axis {a}"
    st = BlockState(sol, np.nan_to_num(prims, nan=1.0), np.nan_to_num(cons, nan=1.0))
    dt = port.time_step_size(prims, s)
    for _ in range(3):
        prims, cons, dt = port.step(prims, cons, dt, s)
        st.step()
    m = H.defined_mask(s)
    assert H.rel_linf(P.host(st.primitives)[:, m], prims[:, m]) <= 1e-11
    assert abs(st.dt.item() - dt) <= 1e-11 * dt


@pytest.mark.parametrize("force_rows", [False, True])
@pytest.mark.parametrize("fs,stencil", [("ROE", "WENO5-Z"), ("CLLF", "WENO6-CU"), ("LLF", "TENO5")])
def test_flux_splitting_3d_all_kernels(fs, stencil, force_rows, monkeypatch):
    """convective_solver = FLUX-SPLITTING through every sweep kernel (march x / y, rows (TMA) or contig z, fused
    epilogue): per-axis rhs and 3 RK3 steps against the oracle."""
    from jaxfluids_b200.engine import BlockState
    if force_rows:
        monkeypatch.setenv("JXF_FORCE_ROWS", "1")
    s = H.make_setup((18, 16, 40), bc="PERIODIC", stencil=stencil)
    s.convective_solver, s.flux_splitting = "FLUX-SPLITTING", fs
    prims, cons = port.initialize(H.smooth_ic(s, seed=23, amp=0.15), s)
    sol = P.make_solver(s)
    p = P.dev(np.nan_to_num(prims, nan=1.0))
    scales = H.rhs_scales(prims, s)
    for a in s.active:
        rhs = sol.new_rhs()
        sol.sweep(a, p, rhs, accumulate=False)
        assert H.rel_linf(P.host(rhs), port.rhs_axis(prims, a, s), scale=scales) <= H.TOL_RHS, f"axis {a}"
    st = BlockState(sol, np.nan_to_num(prims, nan=1.0), np.nan_to_num(cons, nan=1.0))
    dt = port.time_step_size(prims, s)
    for _ in range(3):
        prims, cons, dt = port.step(prims, cons, dt, s)
        st.step()
    m = H.defined_mask(s)
    assert H.rel_linf(P.host(st.primitives)[:, m], prims[:, m]) <= 1e-11
    assert H.rel_linf(P.host(st.conservatives)[:, m], cons[:, m]) <= 1e-11
    assert abs(st.dt.item() - dt) <= 1e-11 * dt


@pytest.mark.parametrize("name", ["generic/lax100_fs_roe_weno6cu_rk3", "generic/woodward200_fs_roe_weno5z_rk3"])
def test_public_api_runs_the_shipped_flux_splitting_examples(name):
    """The reference's Lax and Woodward-Colella example files (shrunk) through InputManager / InitializationManager /
    SimulationManager.simulate, against what the reference produced."""
    from jaxfluids_b200 import InputManager, InitializationManager, SimulationManager
    g, case, num = H.load_golden(name)
    n = len(g["dt"])
    case, num = json.loads(json.dumps(case)), json.loads(json.dumps(num))
    case["general"]["end_step"] = n
    case["general"]["end_time"] = 1e300
    num.setdefault("output", {}).setdefault("logging", {})["level"] = "NONE"
    im = InputManager(case, num)
    assert im.numerical_setup.conservatives.convective_fluxes.convective_solver == "FLUX-SPLITTING"
    buffers = InitializationManager(im).initialization()
    sim = SimulationManager(im)
    sim.simulate(buffers)
    out = sim.final_buffers
    s = H.setup_from_json(case, num)
    m = H.face_halo_mask(s)
    pr = P.host(out.simulation_buffers.material_fields.primitives)
    assert out.time_control_variables.simulation_step == n
    assert H.rel_linf(pr[:, m], g[f"prims_n{n}"][:, m]) <= H.TOL_PRIMS_100


@pytest.mark.parametrize("recon,frozen,stencil,riemann", [
    ("CONSERVATIVE", "ARITHMETIC", "WENO5-Z", "HLLC"), ("CHAR-CONSERVATIVE", "ARITHMETIC", "WENO5-Z", "HLLC"),
    ("CHAR-CONSERVATIVE", "ROE", "TENO5", "HLL"), ("CHAR-PRIMITIVE", "ROE", "WENO5-Z", "HLLC"),
    ("CONSERVATIVE", "ARITHMETIC", "VANLEER", "AUSMP"), ("FLUX-SPLITTING", "ROE", "WENO5-JS", "HLLC")])
def test_reconstruction_variables_and_roe_frozen_state(recon, frozen, stencil, riemann):
    """reconstruction_variable CONSERVATIVE / CHAR-CONSERVATIVE and frozen_state ROE (godunov and flux_splitting blocks)
    through the generic kernel instantiations in 3-D: per-axis rhs and 3 steps against the oracle."""
    from jaxfluids_b200.engine import BlockState
    fs = recon == "FLUX-SPLITTING"
    s = H.make_setup((16, 12, 40), bc="SYMMETRY", stencil=stencil, recon="CHAR-PRIMITIVE" if fs else recon, riemann=riemann)
    s.frozen_state = frozen
    if fs:
        s.convective_solver, s.flux_splitting = "FLUX-SPLITTING", "CLLF"
    prims, cons = port.initialize(H.smooth_ic(s, seed=29, amp=0.15), s)
    sol = P.make_solver(s)
    p = P.dev(np.nan_to_num(prims, nan=1.0))
    scales = H.rhs_scales(prims, s)
    for a in s.active:
        rhs = sol.new_rhs()
        sol.sweep(a, p, rhs, accumulate=False)
        assert H.rel_linf(P.host(rhs), port.rhs_axis(prims, a, s), scale=scales) <= H.TOL_RHS, f"axis {a}"
    st = BlockState(sol, np.nan_to_num(prims, nan=1.0), np.nan_to_num(cons, nan=1.0))
    dt = port.time_step_size(prims, s)
    for _ in range(3):
        prims, cons, dt = port.step(prims, cons, dt, s)
        st.step()
    m = H.defined_mask(s)
    assert H.rel_linf(P.host(st.primitives)[:, m], prims[:, m]) <= 1e-11
    assert abs(st.dt.item() - dt) <= 1e-11 * dt


@pytest.mark.parametrize("name", H.api_golden_names())
def test_public_api_runs_space_dependent_dirichlet_example(name):
    """The reference's 2-D heat equation example (shrunk): DIRICHLET data given as a lambda of the transverse
    coordinate -- halo slabs written by the host runtime after every halo fill (tests/test_dirichlet_cpu.py checks
    that logic on CPU tensors) -- through InputManager / InitializationManager / SimulationManager.simulate."""
    from jaxfluids_b200 import InputManager, InitializationManager, SimulationManager
    g, case, num = H.load_golden(name)
    n = len(g["dt"])
    case, num = json.loads(json.dumps(case)), json.loads(json.dumps(num))
    case["general"]["end_step"] = n
    case["general"]["end_time"] = 1e300
    num.setdefault("output", {}).setdefault("logging", {})["level"] = "NONE"
    im = InputManager(case, num)
    buffers = InitializationManager(im).initialization()
    s = H.setup_from_json(case, num)
    m = H.defined_mask(s)
    p0 = P.host(buffers.simulation_buffers.material_fields.primitives)
    assert np.array_equal(p0[:, m], g["prims0_halo"][:, m])          # initial halos incl. the lambda's values, bit-exact
    assert abs(buffers.time_control_variables.physical_timestep_size - float(g["dt0"])) <= 1e-14 * float(g["dt0"])
    sim = SimulationManager(im)
    assert sim.runtime.face_data            # boundary data applied in the kernels (jxf_set_face_data)
    sim.simulate(buffers)
    out = sim.final_buffers
    pr = P.host(out.simulation_buffers.material_fields.primitives)
    assert out.time_control_variables.simulation_step == n
    assert H.rel_linf(pr[:, m], g[f"prims_n{n}"][:, m]) <= H.TOL_PRIMS_100
    assert abs(out.time_control_variables.physical_timestep_size - g["dt"][n - 1]) <= 1e-10 * g["dt"][n - 1]
