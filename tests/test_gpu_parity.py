"""GPU parity: the sm_100a kernels, called through the C ABI, against the CPU oracle
(oracle/port.py, pinned bit-identically to the reference) and the reference's own fixtures.

Tolerances are the north-star's: per-stage RHS rel L-inf <= 1e-12, primitives after 100 steps
<= 1e-9, conserved totals <= 1e-12.
"""
import os

import numpy as np
import pytest
import torch

from oracle import port
from tests import helpers as H

pytestmark = pytest.mark.gpu


def make_solver(s: port.Setup, bc=None):
    from jaxfluids_b200.engine import BlockConfig, BlockSolver
    cfg = BlockConfig(cells=s.cells, inv_dx=tuple(float(x) for x in s.inv_dx), dx_min=float(s.dx_min),
                      gamma=s.gamma, bc=bc or s.bc, nh=s.nh, recon=s.recon, stencil=s.stencil, riemann=s.riemann,
                      signal_speed=s.signal_speed, convective_solver=s.convective_solver, flux_splitting=s.flux_splitting,
                      frozen_state=s.frozen_state,
                      integrator=s.integrator, cfl=s.cfl,
                      is_viscous_flux=s.is_viscous_flux, is_heat_flux=s.is_heat_flux,
                      is_viscous_heat_production=s.is_viscous_heat_production, dynamic_viscosity=s.dynamic_viscosity,
                      bulk_viscosity=s.bulk_viscosity, thermal_conductivity_model=s.thermal_conductivity_model,
                      thermal_conductivity=s.thermal_conductivity, prandtl_number=s.prandtl_number,
                      gas_constant=s.gas_constant, is_interpolation_limiter=s.is_interpolation_limiter,
                      limit_velocity=s.limit_velocity, flux_limiter=s.flux_limiter, flux_partition=s.flux_partition,
                      wall_velocity=dict(s.wall_velocity),
                      dirichlet=dict(s.dirichlet), is_volume_force=s.is_volume_force, gravity=tuple(s.gravity),
                      is_convective_flux=s.is_convective_flux)
    return BlockSolver(cfg)


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).cuda()


def host(t):
    return t.detach().cpu().numpy()


def test_library_is_loaded_and_native():
    from jaxfluids_b200 import _lib
    lib = _lib.load()
    assert lib.jxf_version() >= 100


VARIANTS = [("CHAR-PRIMITIVE", "HLLC"), ("PRIMITIVE", "HLLC"), ("CHAR-PRIMITIVE", "RUSANOV"), ("PRIMITIVE", "RUSANOV")]
GRIDS = [(48, 1, 1), (1000, 1, 1), (40, 24, 1), (33, 47, 1), (24, 20, 28), (16, 33, 40), (37, 16, 19)]


@pytest.mark.parametrize("recon,riemann", VARIANTS)
@pytest.mark.parametrize("cells", GRIDS)
def test_rhs_per_axis_and_total(cells, recon, riemann):
    s = H.make_setup(cells, bc="PERIODIC", recon=recon, riemann=riemann)
    prims, cons = port.initialize(H.smooth_ic(s, seed=sum(cells)), s)
    sol = make_solver(s)
    p = dev(prims)
    ref = port.compute_rhs(prims, s)
    scales = H.rhs_scales(prims, s)
    for a in s.active:
        rhs = sol.new_rhs()
        sol.sweep(a, p, rhs, accumulate=False)
        assert H.rel_linf(host(rhs), port.rhs_axis(prims, a, s), scale=scales) <= H.TOL_RHS, f"axis {a}"
    got = host(sol.compute_rhs(p))
    assert H.rel_linf(got, ref, scale=scales) <= H.TOL_RHS


@pytest.mark.parametrize("name", H.golden_names())
def test_reference_fixture_rhs_and_stages(name):
    """Against what the reference itself produced (tests/golden): per-axis rhs, stage rhs, stage states."""
    g, case, num = H.load_golden(name)
    s = H.setup_from_json(case, num)
    sol = make_solver(s)
    p0 = dev(g["prims0_halo"])
    scales = H.rhs_scales(g["prims0_halo"], s)
    if s.flux_limiter:          # the fixture's rhs were taken with the step's physical time step size
        sol.bind_timestep(dev(np.array([float(g["dt0"])])))
    for a in s.active:
        rhs = sol.new_rhs()
        sol.sweep(a, p0, rhs, accumulate=False)
        assert H.rel_linf(host(rhs), g[f"rhs_axis{a}"], scale=scales) <= H.TOL_RHS
    mask = H.face_halo_mask(s)
    nst = port.RK[s.integrator]["stages"]
    for k in range(nst):
        # stage rhs from the reference's stage-entry primitives
        p_in = g["prims0_halo"] if k == 0 else g[f"prims_s{k-1}"]
        p_in = np.nan_to_num(p_in, nan=1.0, posinf=1.0, neginf=1.0)
        got = host(sol.compute_rhs(dev(p_in)))
        assert H.rel_linf(got, g[f"rhs_s{k}"], scale=H.rhs_scales(p_in, s)) <= H.TOL_RHS, f"stage {k}"
    # full first step through the fused stage kernels
    from jaxfluids_b200.engine import BlockState
    st = BlockState(sol, np.nan_to_num(g["prims0_halo"], nan=1.0, posinf=1.0, neginf=1.0),
                    np.nan_to_num(g["cons0_halo"], nan=1.0, posinf=1.0, neginf=1.0))
    assert abs(st.dt.item() - float(g["dt0"])) <= 1e-14 * float(g["dt0"])
    st.step()
    pr, co = host(st.primitives), host(st.conservatives)
    ref_p, ref_c = g[f"prims_s{nst-1}"], g[f"cons_s{nst-1}"]
    assert H.rel_linf(pr[:, mask], ref_p[:, mask]) <= 1e-12
    assert H.rel_linf(co[:, mask], ref_c[:, mask]) <= 1e-12
    assert abs(st.dt.item() - g["dt"][0]) <= 1e-12 * g["dt"][0]


@pytest.mark.parametrize("name", H.golden_names())
def test_reference_fixture_multi_step(name):
    """dt sequence, totals, min rho / min p and the state after N steps against the reference."""
    from jaxfluids_b200.engine import BlockState
    g, case, num = H.load_golden(name)
    s = H.setup_from_json(case, num)
    sol = make_solver(s)
    st = BlockState(sol, np.nan_to_num(g["prims0_halo"], nan=1.0, posinf=1.0, neginf=1.0),
                    np.nan_to_num(g["cons0_halo"], nan=1.0, posinf=1.0, neginf=1.0))
    mask = H.face_halo_mask(s)
    n = len(g["dt"])
    sl = (slice(None),) + s.interior
    for i in range(1, n + 1):
        st.step()
        info = host(st.info)
        assert abs(st.dt.item() - g["dt"][i - 1]) <= 1e-10 * g["dt"][i - 1]
        assert abs(st.time.item() - g["time"][i - 1]) <= 1e-12 * g["time"][i - 1]
        assert abs(info[1] - g["min_density"][i - 1]) <= 1e-10 * abs(g["min_density"][i - 1])
        assert abs(info[2] - g["min_pressure"][i - 1]) <= 1e-10 * abs(g["min_pressure"][i - 1])
        if f"prims_n{i}" in g:
            pr, co = host(st.primitives), host(st.conservatives)
            assert H.rel_linf(pr[:, mask], g[f"prims_n{i}"][:, mask]) <= H.TOL_PRIMS_100, f"step {i}"
            tot = np.array([co[sl][v].sum() for v in range(5)])
            ref = g["totals"][i - 1]
            scale = np.maximum(np.abs(ref), 1e-3 * np.max(np.abs(ref)))
            assert np.max(np.abs(tot - ref) / scale) <= H.TOL_TOTALS * max(1.0, np.sqrt(np.prod(s.cells)) / 10)


@pytest.mark.parametrize("bc", ["PERIODIC", "SYMMETRY", "ZEROGRADIENT"])
@pytest.mark.parametrize("cells", [(32, 1, 1), (20, 24, 1), (12, 16, 20)])
def test_halo_fill(cells, bc):
    s = H.make_setup(cells, bc=bc)
    sol = make_solver(s)
    rng = np.random.default_rng(1)
    base = np.ones(s.shape) * port.EPS
    base[(slice(None),) + s.interior] = H.smooth_ic(s, seed=3)
    cons = port.cons_from_prims(base, s.gamma)
    ref_p, ref_c = port.halo_fill(base, cons, s)
    p, c = dev(base), dev(cons)
    sol.halo_fill(p, c)
    assert np.array_equal(host(p), ref_p)            # pure copies / sign flips: bit exact
    mask = H.face_halo_mask(s)
    assert H.rel_linf(host(c)[:, mask], ref_c[:, mask]) <= 1e-15


@pytest.mark.parametrize("cells", [(64, 1, 1), (24, 20, 1), (12, 16, 20)])
def test_whole_buffer_transforms(cells):
    s = H.make_setup(cells)
    sol = make_solver(s)
    prims, cons = port.initialize(H.smooth_ic(s, seed=5), s)
    prims = np.nan_to_num(prims)
    prims[prims == 0] = 1.0
    c = sol.new_field()
    sol.cons_from_prims(dev(prims), c)
    ref_c = port.cons_from_prims(prims, s.gamma)
    assert H.rel_linf(host(c), ref_c) <= 1e-15
    p = sol.new_field()
    sol.prims_from_cons(dev(ref_c), p)
    assert H.rel_linf(host(p), port.prims_from_cons(ref_c, s.gamma)) <= 1e-14


@pytest.mark.parametrize("integrator", ["EULER", "RK2", "RK3"])
@pytest.mark.parametrize("cells,bc", [((96, 1, 1), "ZEROGRADIENT"), ((28, 36, 1), "SYMMETRY"), ((16, 12, 20), "PERIODIC")])
def test_steps_against_oracle(cells, bc, integrator):
    """10 steps: state, dt sequence, min rho/p, totals."""
    from jaxfluids_b200.engine import BlockState
    s = H.make_setup(cells, bc=bc, integrator=integrator)
    prims, cons = port.initialize(H.smooth_ic(s, seed=11, amp=0.1), s)
    sol = make_solver(s)
    st = BlockState(sol, np.nan_to_num(prims, nan=1.0), np.nan_to_num(cons, nan=1.0))
    dt = port.time_step_size(prims, s)
    assert abs(st.dt.item() - dt) <= 1e-14 * dt
    mask = H.face_halo_mask(s)
    for i in range(10):
        prims, cons, dt = port.step(prims, cons, dt, s)
        st.step()
    assert abs(st.dt.item() - dt) <= 1e-11 * dt
    assert H.rel_linf(host(st.primitives)[:, mask], prims[:, mask]) <= 1e-11
    assert H.rel_linf(host(st.conservatives)[:, mask], cons[:, mask]) <= 1e-11
    mr, mp = port.positivity_info(prims, s)
    info = host(st.info)
    assert abs(info[1] - mr) <= 1e-11 * abs(mr) and abs(info[2] - mp) <= 1e-11 * abs(mp)


def test_sod_100_steps_within_1e9():
    """North-star bound: primitives rel L-inf <= 1e-9 after 100 steps (Sod, 1000 cells)."""
    from jaxfluids_b200.engine import BlockState
    s = H.make_setup((1000, 1, 1), bc="ZEROGRADIENT")
    x = s.cell_centers()[0]
    pr = np.zeros((5, 1000, 1, 1))
    pr[0, :, 0, 0] = np.where(x <= 0.5, 1.0, 0.125)
    pr[4, :, 0, 0] = np.where(x <= 0.5, 1.0, 0.1)
    prims, cons = port.initialize(pr, s)
    sol = make_solver(s)
    st = BlockState(sol, prims, cons)
    dt = port.time_step_size(prims, s)
    for i in range(100):
        prims, cons, dt = port.step(prims, cons, dt, s)
        st.step()
    assert H.rel_linf(host(st.primitives), prims) <= H.TOL_PRIMS_100
    tot = host(st.conservatives)[(slice(None),) + s.interior].sum(axis=(1, 2, 3))
    ref = port.totals(cons, s)
    assert abs(tot[0] - ref[0]) <= H.TOL_TOTALS * abs(ref[0]) * 10
    assert abs(tot[4] - ref[4]) <= H.TOL_TOTALS * abs(ref[4]) * 10


def test_tgv_32_100_steps_within_1e9():
    """TGV 32^3 (the shipped case file's grid), SYMMETRY, 100 steps vs the oracle."""
    from jaxfluids_b200.engine import BlockState
    g, case, num = H.load_golden("tgv16_sym_char_hllc_rk3")
    case["domain"]["x"]["cells"] = case["domain"]["y"]["cells"] = case["domain"]["z"]["cells"] = 32
    s = H.setup_from_json(case, num)
    x, y, z = np.meshgrid(*s.cell_centers(), indexing="ij")
    pr = np.stack([np.ones_like(x), np.sin(x) * np.cos(y) * np.cos(z), -np.cos(x) * np.sin(y) * np.cos(z),
                   np.zeros_like(x),
                   1 / 1.4 / 0.1 ** 2 + 1 / 16.0 * ((np.cos(2 * x) + np.cos(2 * y)) * (np.cos(2 * z) + 2))])
    prims, cons = port.initialize(pr, s)
    sol = make_solver(s)
    st = BlockState(sol, np.nan_to_num(prims, nan=1.0), np.nan_to_num(cons, nan=1.0))
    dt = port.time_step_size(prims, s)
    mask = H.face_halo_mask(s)
    for i in range(100):
        prims, cons, dt = port.step(prims, cons, dt, s)
        st.step()
    assert H.rel_linf(host(st.primitives)[:, mask], prims[:, mask]) <= H.TOL_PRIMS_100
    tot = host(st.conservatives)[(slice(None),) + s.interior].sum(axis=(1, 2, 3))
    ref = port.totals(cons, s)
    assert abs(tot[0] - ref[0]) <= H.TOL_TOTALS * abs(ref[0]) * 10
    assert abs(tot[4] - ref[4]) <= H.TOL_TOTALS * abs(ref[4]) * 10


def test_pack_unpack_faces_roundtrip():
    """pack(face) of a periodic block + unpack on the opposite face == the PERIODIC halo fill."""
    s = H.make_setup((12, 16, 20), bc="PERIODIC")
    nb = {f: "NEIGHBOR" for f in port.FACES}
    sol_n = make_solver(s, bc=nb)
    base = np.ones(s.shape) * port.EPS
    base[(slice(None),) + s.interior] = H.smooth_ic(s, seed=2)
    cons = port.cons_from_prims(base, s.gamma)
    ref_p, ref_c = port.halo_fill(base, cons, s)
    p, c = dev(base), dev(cons)
    sol_n.halo_fill(p, c)                       # all faces NEIGHBOR: must be a no-op
    assert np.array_equal(host(p), base)
    opp = {0: 1, 1: 0, 2: 3, 3: 2, 4: 5, 5: 4}
    for face in range(6):
        slab = torch.empty(sol_n.face_slab_elems(face), dtype=torch.float64, device="cuda")
        sol_n.pack_face(face, p, slab)
        sol_n.unpack_face(opp[face], slab, p, c)
    assert np.array_equal(host(p), ref_p)
    mask = H.face_halo_mask(s)
    assert H.rel_linf(host(c)[:, mask], ref_c[:, mask]) <= 1e-15


def test_full_size_properties_tgv256():
    """BASELINE config 3 size (TGV 256^3): size-independent properties instead of the oracle --
    (i) mass and energy conserved to 1e-12 under SYMMETRY BCs, (ii) the TGV mirror symmetries hold."""
    from jaxfluids_b200.engine import BlockState
    n = 256
    s = H.make_setup((n, n, n), bc="SYMMETRY", gamma=5.0 / 3.0, length=2 * np.pi)
    c = torch.as_tensor(s.cell_centers()[0], dtype=torch.float64, device="cuda")
    x, y, z = c[:, None, None], c[None, :, None], c[None, None, :]
    sol = make_solver(s)
    p = sol.new_field(port.EPS)
    nh = s.nh
    it = p[:, nh:-nh, nh:-nh, nh:-nh]
    it[0] = 1.0
    it[1] = torch.sin(x) * torch.cos(y) * torch.cos(z)
    it[2] = -torch.cos(x) * torch.sin(y) * torch.cos(z)
    it[3] = 0.0
    it[4] = 1 / 1.4 / 0.1 ** 2 + 1 / 16.0 * ((torch.cos(2 * x) + torch.cos(2 * y)) * (torch.cos(2 * z) + 2))
    co = sol.new_field(port.EPS)
    sol.cons_from_prims(p, co)
    sol.halo_fill(p, co)
    st = BlockState(sol, p, co)
    tot0 = st.conservatives[:, nh:-nh, nh:-nh, nh:-nh].sum(dim=(1, 2, 3)).cpu().numpy()
    for _ in range(3):
        st.step()
    ci = st.conservatives[:, nh:-nh, nh:-nh, nh:-nh]
    tot = ci.sum(dim=(1, 2, 3)).cpu().numpy()
    assert abs(tot[0] - tot0[0]) <= 1e-12 * abs(tot0[0])
    assert abs(tot[4] - tot0[4]) <= 1e-12 * abs(tot0[4])
    pi = st.primitives[:, nh:-nh, nh:-nh, nh:-nh]
    # mirror symmetry about the domain centre planes: rho, p even; u odd in x; v odd in y
    assert float((pi[0] - pi[0].flip(0)).abs().max()) <= 1e-12
    assert float((pi[1] + pi[1].flip(0)).abs().max()) <= 1e-12
    assert float((pi[2] + pi[2].flip(1)).abs().max()) <= 1e-12
    assert float((pi[4] - pi[4].flip(2)).abs().max()) <= 1e-11 * float(pi[4].abs().max())
    assert bool(torch.isfinite(pi).all())


def test_fast_reciprocal_and_rsqrt_accuracy():
    """rcp_fast / rsqrt_fast (MUFU seed + Newton) are accurate to a few ulp over 40 decades."""
    import ctypes as C
    from jaxfluids_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    x = np.concatenate([10.0 ** rng.uniform(-200, 200, 200000), rng.uniform(0.5, 2.0, 200000),
                        10.0 ** rng.uniform(-30, 12, 200000)])
    xd = dev(x)
    out = torch.empty((x.size, 4), dtype=torch.float64, device="cuda")
    _lib.check(lib.jxf_debug_math(C.c_void_p(xd.data_ptr()), x.size, C.c_void_p(out.data_ptr()), None))
    o = host(out)
    seed_rcp = np.max(np.abs(o[:, 0] * x - 1.0))
    seed_rsq = np.max(np.abs(o[:, 1] * o[:, 1] * x - 1.0))
    err_rcp = np.max(np.abs(o[:, 2] * x - 1.0))
    err_rsq = np.max(np.abs(o[:, 3] * np.sqrt(x) - 1.0))
    print(f"seed rcp {seed_rcp:.3e} seed rsqrt(y^2 x - 1) {seed_rsq:.3e} rcp_fast {err_rcp:.3e} rsqrt_fast {err_rsq:.3e}")
    assert seed_rcp < 2.0 ** -18 and seed_rsq < 2.0 ** -17
    assert err_rcp <= 4.5e-16 and err_rsq <= 4.5e-16


@pytest.mark.parametrize("no_tma", [0, 1])
@pytest.mark.parametrize("cells,bc", [((12, 16, 40), "PERIODIC"), ((8, 9, 70), "SYMMETRY"), ((10, 12, 33), "PERIODIC"),
                                      ((6, 40, 64), "SYMMETRY"), ((24, 64, 1), "ZEROGRADIENT"), ((96, 1, 1), "ZEROGRADIENT")])
def test_rows_kernel_tma_and_cp_async(cells, bc, no_tma, monkeypatch):
    """The production contiguous-axis kernel (32-row groups, windows staged by TMA or cp.async) forced
    onto small grids: rhs, and 3 full steps incl. fused epilogue/halo images, against the oracle."""
    from jaxfluids_b200.engine import BlockState
    monkeypatch.setenv("JXF_FORCE_ROWS", "1")
    monkeypatch.setenv("JXF_NO_TMA", str(no_tma))
    s = H.make_setup(cells, bc=bc)
    prims, cons = port.initialize(H.smooth_ic(s, seed=7, amp=0.1), s)
    sol = make_solver(s)
    prims_c = np.nan_to_num(prims, nan=1.0)
    got = host(sol.compute_rhs(dev(prims_c)))
    assert H.rel_linf(got, port.compute_rhs(prims, s), scale=H.rhs_scales(prims, s)) <= H.TOL_RHS
    st = BlockState(sol, prims_c, np.nan_to_num(cons, nan=1.0))
    dt = port.time_step_size(prims, s)
    for _ in range(3):
        prims, cons, dt = port.step(prims, cons, dt, s)
        st.step()
    mask = H.face_halo_mask(s)
    assert H.rel_linf(host(st.primitives)[:, mask], prims[:, mask]) <= 1e-12
    assert H.rel_linf(host(st.conservatives)[:, mask], cons[:, mask]) <= 1e-12
    assert abs(st.dt.item() - dt) <= 1e-12 * dt


@pytest.mark.parametrize("no_march,order", [(0, "xzy"), (0, "xyz")])
@pytest.mark.parametrize("cells,bc,recon,riemann", [
    ((20, 16, 12), "PERIODIC", "CHAR-PRIMITIVE", "HLLC"), ((12, 36, 10), "SYMMETRY", "PRIMITIVE", "HLLC"),
    ((40, 8, 9), "ZEROGRADIENT", "CHAR-PRIMITIVE", "RUSANOV"), ((33, 20, 1), "SYMMETRY", "CHAR-PRIMITIVE", "HLLC")])
def test_strided_forms_and_sweep_orders(cells, bc, recon, riemann, no_march, order, monkeypatch):
    """The marching sweep (shared-memory ring `sweep_march`; its register-window predecessor `sweep_strided` is
    compiled only with -DJXF_WITH_STRIDED) and both stage sweep orders (epilogue on the marching y sweep, or on the contiguous z sweep): rhs and
    3 full steps incl. fused epilogue / halo images, against the oracle."""
    from jaxfluids_b200.engine import BlockState
    monkeypatch.setenv("JXF_NO_MARCH", str(no_march))
    monkeypatch.setenv("JXF_SWEEP_ORDER", order)
    s = H.make_setup(cells, bc=bc, recon=recon, riemann=riemann)
    prims, cons = port.initialize(H.smooth_ic(s, seed=11, amp=0.1), s)
    sol = make_solver(s)
    prims_c = np.nan_to_num(prims, nan=1.0)
    got = host(sol.compute_rhs(dev(prims_c)))
    assert H.rel_linf(got, port.compute_rhs(prims, s), scale=H.rhs_scales(prims, s)) <= H.TOL_RHS
    st = BlockState(sol, prims_c, np.nan_to_num(cons, nan=1.0))
    dt = port.time_step_size(prims, s)
    for _ in range(3):
        prims, cons, dt = port.step(prims, cons, dt, s)
        st.step()
    mask = H.face_halo_mask(s)
    assert H.rel_linf(host(st.primitives)[:, mask], prims[:, mask]) <= 1e-12
    assert H.rel_linf(host(st.conservatives)[:, mask], cons[:, mask]) <= 1e-12
    assert abs(st.dt.item() - dt) <= 1e-12 * dt


@pytest.mark.parametrize("cells,bc,integrator", [((20, 12, 16), "SYMMETRY", "RK3"), ((64, 1, 1), "ZEROGRADIENT", "RK2"),
                                                 ((18, 22, 1), "PERIODIC", "EULER")])
def test_step_from_separate_api_pieces(cells, bc, integrator):
    """One full RK step driven piece by piece, as the reference's do_runge_kutta_stages does
    (simulation_manager.py:796-963): compute_rhs -> perform_stage_integration (stand-alone
    jxf_integrate_stage) -> get_primitives_from_conservatives -> halo update, vs the oracle stage by stage."""
    s = H.make_setup(cells, bc=bc, integrator=integrator)
    prims, cons = port.initialize(H.smooth_ic(s, seed=3, amp=0.1), s)
    sol = make_solver(s)
    dt = port.time_step_size(prims, s)
    p, c = dev(np.nan_to_num(prims, nan=1.0)), dev(np.nan_to_num(cons, nan=1.0))
    c_n = c.clone()
    cons_n = cons
    mask = H.face_halo_mask(s)
    for k in range(sol.stages):
        rhs = sol.compute_rhs(p)
        c_new = sol.integrate_stage(k, c, c_n if k > 0 else None, rhs, dt)
        assert c_new.data_ptr() != c.data_ptr()                 # functional: new buffer, like the reference
        p_new = torch.empty_like(c_new)
        sol.prims_from_cons(c_new, p_new)
        sol.halo_fill(p_new, c_new)
        with np.errstate(all="ignore"):
            prims, cons, rhs_ref = port.stage(prims, cons, cons_n, dt, k, s)
        assert H.rel_linf(host(rhs), rhs_ref, scale=H.rhs_scales(np.nan_to_num(host(p), nan=1.0), s)) <= H.TOL_RHS
        assert H.rel_linf(host(c_new)[:, mask], cons[:, mask]) <= 1e-13, f"stage {k}"
        assert H.rel_linf(host(p_new)[:, mask], prims[:, mask]) <= 1e-13, f"stage {k}"
        p, c = p_new, c_new
    # in-place form (cons_out aliases cons) gives the same bits
    c2 = c_n.clone()
    rhs0 = sol.compute_rhs(dev(np.nan_to_num(port.initialize(H.smooth_ic(s, seed=3, amp=0.1), s)[0], nan=1.0)))
    a = sol.integrate_stage(0, c_n, None, rhs0, dt)
    sol.integrate_stage(0, c2, None, rhs0, dt, out=c2)
    assert torch.equal(a, c2)


# ---------------------------------------------------------------------------
# viscous + heat flux (CENTRAL4), edge halos, diffusive dt limits
# ---------------------------------------------------------------------------
DISS = [dict(is_viscous_flux=True, dynamic_viscosity=2e-3, bulk_viscosity=5e-4),
        dict(is_viscous_flux=True, is_heat_flux=True, dynamic_viscosity=1e-3, thermal_conductivity_model="PRANDTL",
             prandtl_number=0.71, gas_constant=1.3),
        dict(is_heat_flux=True, thermal_conductivity=3e-3, gas_constant=0.8),
        dict(is_viscous_flux=True, is_viscous_heat_production=False, dynamic_viscosity=1e-3)]


def diss_setup(cells, bc, extra, **kw):
    s = H.make_setup(cells, bc=bc, **kw)
    for k, v in extra.items():
        setattr(s, k, v)
    return s


@pytest.mark.parametrize("bc", ["PERIODIC", "SYMMETRY", "ZEROGRADIENT"])
@pytest.mark.parametrize("cells", [(14, 12, 16), (18, 22, 1), (40, 1, 1)])
def test_edge_halo_fill_is_bit_exact(cells, bc):
    """jxf_halo_fill with the dissipative fluxes on = face halos + edge halos
    (halos/outer/material.py:289-383): bit-identical to the oracle on every face and edge cell."""
    s = diss_setup(cells, bc, DISS[0])
    prims, cons = port.initialize(H.smooth_ic(s, seed=5), s)            # oracle: faces + edges
    s0 = H.make_setup(cells, bc=bc)
    p0, c0 = port.initialize(H.smooth_ic(s, seed=5), s0)                # faces only
    sol = make_solver(s)
    p, c = dev(np.nan_to_num(p0, nan=7.0)), dev(np.nan_to_num(c0, nan=7.0))
    if len(s.active) > 1:                                               # scramble the edge regions first
        nh = s.nh
        p[:, :nh, :nh] = 3.0
    sol.halo_fill(p, c)
    m = H.defined_mask(s)                                               # interior + faces + edges
    assert np.array_equal(host(p)[:, m], prims[:, m])              # copies / sign flips / 0.5 (a + b): bit exact
    assert H.rel_linf(host(c)[:, m], cons[:, m]) <= 1e-15           # recomputed (FMA contraction on the device)


@pytest.mark.parametrize("extra", DISS)
@pytest.mark.parametrize("cells,bc", [((20, 16, 24), "PERIODIC"), ((12, 18, 40), "SYMMETRY"), ((16, 12, 10), "ZEROGRADIENT"),
                                      ((30, 26, 1), "SYMMETRY"), ((24, 33, 1), "ZEROGRADIENT"), ((64, 1, 1), "ZEROGRADIENT"),
                                      ((6, 8, 600), "PERIODIC"), ((8, 1100, 1), "SYMMETRY")])
def test_dissipative_rhs_per_axis_and_total(cells, bc, extra):
    """Per-axis and total rhs with the viscous / heat flux folded in (space_solver.py:567-599), and the
    stand-alone dissipative sweep against (oracle with) - (oracle without)."""
    s = diss_setup(cells, bc, extra)
    s_conv = H.make_setup(cells, bc=bc)
    prims, cons = port.initialize(H.smooth_ic(s, seed=sum(cells), amp=0.1), s)
    sol = make_solver(s)
    p = dev(np.nan_to_num(prims, nan=1.0))
    scales = H.rhs_scales(prims, s_conv)
    for a in s.active:
        rhs = sol.new_rhs()
        sol.sweep(a, p, rhs, accumulate=False)
        assert H.rel_linf(host(rhs), port.rhs_axis(prims, a, s), scale=scales) <= H.TOL_RHS, f"axis {a}"
        d = sol.new_rhs()
        sol.dissipative_sweep(a, p, d, accumulate=False)
        ref_d = port.rhs_axis(prims, a, s) - port.rhs_axis(prims, a, s_conv)
        dscale = np.maximum(np.abs(ref_d).max(axis=(1, 2, 3)), 1e-300)
        err = np.abs(host(d) - ref_d).max(axis=(1, 2, 3))
        # the reference forms (conv - visc) before differencing: its own rounding is ~1e-16 of the CONVECTIVE terms
        assert np.all(err <= 1e-12 * np.maximum(dscale, scales.reshape(-1))), (a, err, dscale)
        assert np.all(host(d)[0] == 0.0)
    got = host(sol.compute_rhs(p))
    assert H.rel_linf(got, port.compute_rhs(prims, s), scale=scales) <= H.TOL_RHS


@pytest.mark.parametrize("extra", DISS[:3])
@pytest.mark.parametrize("cells,bc,integrator", [((16, 12, 20), "SYMMETRY", "RK3"), ((14, 16, 12), "PERIODIC", "RK2"),
                                                 ((22, 18, 1), "ZEROGRADIENT", "RK3"), ((48, 1, 1), "ZEROGRADIENT", "EULER")])
def test_dissipative_steps_against_oracle(cells, bc, integrator, extra):
    """5 full steps with viscous / heat flux: state, dt sequence (incl. the diffusive dt limits) and the
    temperature buffer."""
    from jaxfluids_b200.engine import BlockState
    extra = dict(extra)
    if extra.get("dynamic_viscosity"):
        extra["dynamic_viscosity"] = 0.05                                # large enough for the viscous dt limit to bind
    s = diss_setup(cells, bc, extra, integrator=integrator)
    prims, cons = port.initialize(H.smooth_ic(s, seed=9, amp=0.1), s)
    sol = make_solver(s)
    st = BlockState(sol, np.nan_to_num(prims, nan=1.0), np.nan_to_num(cons, nan=1.0))
    dt = port.time_step_size(prims, s)
    assert abs(st.dt.item() - dt) <= 1e-14 * dt
    conv_dt = port.time_step_size(prims, H.make_setup(cells, bc=bc, integrator=integrator))
    assert dt <= conv_dt                                                 # the diffusive limits can only shorten it
    m = H.defined_mask(s)
    for _ in range(5):
        prims, cons, dt = port.step(prims, cons, dt, s)
        st.step()
        assert abs(st.dt.item() - dt) <= 1e-12 * dt
    assert H.rel_linf(host(st.primitives)[:, m], prims[:, m]) <= 1e-12
    assert H.rel_linf(host(st.conservatives)[:, m], cons[:, m]) <= 1e-12
    gp = host(st.primitives)
    T = host(sol.temperature(st.primitives))                          # ideal_gas.py:64-65 on the same bits: exact
    assert np.array_equal(T[m], port.temperature(gp, s)[m])
    assert np.allclose(T[m], port.temperature(prims, s)[m], rtol=1e-11, atol=0)


@pytest.mark.parametrize("name", ["tgv12_sym_visc_prandtl_rk3", "riemann2d_20x24_visc_rk3", "tgv16_sym_char_hllc_rk3",
                                  "cavity_24x20_wall_js_visc_rk3", "sod100_js_char_hllc_rk3", "rti_16x48_dirichlet_gravity_rk3",
                                  "heat1d_40_dirichlet_noconv_rk3"])
def test_public_api_runs_reference_case_files(name):
    """The reference's JSON setups through InputManager -> InitializationManager -> SimulationManager
    (do_integration_step), compared with what the reference itself produced for them: dt sequence, state after
    N steps, the temperature buffer (viscous cases)."""
    import copy
    from jaxfluids_b200 import InputManager, InitializationManager, SimulationManager
    g, case, num = H.load_golden(name)
    s = H.setup_from_json(case, num)
    num = copy.deepcopy(num)
    num.setdefault("output", {}).setdefault("logging", {})["level"] = "NONE"
    im = InputManager(case, num)
    buffers = InitializationManager(im).initialization()          # the case file's own initial condition
    sim = SimulationManager(im)
    assert sim.halo_manager.fill_edge_halos_material == s.is_dissipative
    mf = buffers.simulation_buffers.material_fields
    assert (mf.temperature is not None) == s.is_dissipative
    assert abs(buffers.time_control_variables.physical_timestep_size - float(g["dt0"])) <= 1e-12 * float(g["dt0"])
    m = H.defined_mask(s)
    assert H.rel_linf(host(mf.primitives)[:, m], g["prims0_halo"][:, m]) <= 1e-14
    n = len(g["dt"])
    for i in range(1, n + 1):
        buffers, _ = sim.do_integration_step(buffers)
        tcv = buffers.time_control_variables
        assert abs(tcv.physical_timestep_size - g["dt"][i - 1]) <= 1e-10 * g["dt"][i - 1]
        if f"prims_n{i}" in g:
            mf = buffers.simulation_buffers.material_fields
            assert H.rel_linf(host(mf.primitives)[:, m], g[f"prims_n{i}"][:, m]) <= H.TOL_PRIMS_100
            if s.is_dissipative:
                gp = host(mf.primitives)
                assert np.array_equal(host(mf.temperature)[m], port.temperature(gp, s)[m])
    assert tcv.simulation_step == n


@pytest.mark.parametrize("recon,riemann", VARIANTS)
@pytest.mark.parametrize("cells,bc", [((96, 1, 1), "ZEROGRADIENT"), ((36, 28, 1), "PERIODIC"), ((20, 16, 40), "SYMMETRY")])
def test_weno5js_rhs_and_steps(cells, bc, recon, riemann):
    """godunov.reconstruction_stencil = WENO5-JS (weno/weno5_js.py): per-axis rhs and 3 steps, all kernels
    (marching with carried weights, rows with TMA windows, small-grid contiguous)."""
    from jaxfluids_b200.engine import BlockState
    s = H.make_setup(cells, bc=bc, recon=recon, riemann=riemann, stencil="WENO5-JS")
    prims, cons = port.initialize(H.smooth_ic(s, seed=2, amp=0.15), s)
    sol = make_solver(s)
    p = dev(np.nan_to_num(prims, nan=1.0))
    scales = H.rhs_scales(prims, s)
    for a in s.active:
        rhs = sol.new_rhs()
        sol.sweep(a, p, rhs, accumulate=False)
        assert H.rel_linf(host(rhs), port.rhs_axis(prims, a, s), scale=scales) <= H.TOL_RHS, f"axis {a}"
    st = BlockState(sol, np.nan_to_num(prims, nan=1.0), np.nan_to_num(cons, nan=1.0))
    dt = port.time_step_size(prims, s)
    for _ in range(3):
        prims, cons, dt = port.step(prims, cons, dt, s)
        st.step()
    m = H.defined_mask(s)
    assert H.rel_linf(host(st.primitives)[:, m], prims[:, m]) <= 1e-12
    assert abs(st.dt.item() - dt) <= 1e-12 * dt


@pytest.mark.parametrize("tag", ["lv0", "lv1"])
def test_interpolation_limiter_fixture_rhs(tag):
    """positivity/is_interpolation_limiter (limiter_interpolation.py:77-209) on the reference's fixture where it
    fires thousands of times (near-vacuum regions): rhs against the reference, all kernels."""
    import json, os
    g = np.load(os.path.join(H.GOLDEN, "special", "limiter_riemann2d_20x24.npz"))
    case, num = json.loads(str(g[f"case_json_{tag}"])), json.loads(str(g[f"num_json_{tag}"]))
    s = H.setup_from_json(case, num)
    sol = make_solver(s)
    p = dev(np.nan_to_num(g[f"prims_halo_{tag}"], nan=1.0, posinf=1.0, neginf=1.0))
    ref = g[f"rhs_{tag}"]
    with np.errstate(all="ignore"):
        scales = H.rhs_scales(g[f"prims_halo_{tag}"], s)
    got = host(sol.compute_rhs(p))
    assert np.isfinite(got).all()
    assert H.rel_linf(got, ref, scale=scales) <= H.TOL_RHS
    # and the limiter matters: without it the result differs
    import copy
    s0 = copy.copy(s)
    s0.is_interpolation_limiter = False
    got0 = host(make_solver(s0).compute_rhs(p))
    assert not np.allclose(got0, got, rtol=1e-6, atol=0, equal_nan=True)


@pytest.mark.parametrize("cells,walls", [
    ((28, 24, 1), {"east": (0, 0, 0), "west": (0, 0, 0), "north": (0.5, 0, 0), "south": (0, 0, 0)}),
    ((12, 16, 20), {"north": (0.1, 0.0, -0.2), "south": (0, 0, 0)})])
def test_wall_boundaries_viscous_steps(cells, walls):
    """WALL faces with constant wall velocity (halos/outer/material.py:473-520: u_halo = 2 u_wall - u_mirror), the
    other active faces PERIODIC; viscous; halo fill bit-exact, 4 steps against the oracle (lid-driven-cavity and
    Couette-like setups)."""
    from jaxfluids_b200.engine import BlockState
    bc = {f: ("WALL" if f in walls else "PERIODIC") for f in port.FACES}
    s = H.make_setup(cells, bc=bc, recon="PRIMITIVE", stencil="WENO5-JS", nh=4)
    s.wall_velocity = {f: tuple(float(x) for x in v) for f, v in walls.items()}
    s.is_viscous_flux, s.is_viscous_heat_production, s.dynamic_viscosity = True, False, 5e-3
    s.is_interpolation_limiter = True
    prims, cons = port.initialize(H.smooth_ic(s, seed=13, amp=0.05), s)
    sol = make_solver(s)
    # stand-alone halo fill from interior-only data: copies / 2 u_w - u: bit exact
    s_raw = H.make_setup(cells, bc={f: ("ZEROGRADIENT" if s.bc[f] != "INACTIVE" else "INACTIVE") for f in port.FACES}, nh=4)
    p_raw, c_raw = port.initialize(H.smooth_ic(s, seed=13, amp=0.05), s_raw)
    p, c = dev(p_raw), dev(c_raw)
    sol.halo_fill(p, c)
    m = H.defined_mask(s)
    assert np.array_equal(host(p)[:, m], prims[:, m])
    st = BlockState(sol, np.nan_to_num(prims, nan=1.0), np.nan_to_num(cons, nan=1.0))
    dt = port.time_step_size(prims, s)
    for _ in range(4):
        prims, cons, dt = port.step(prims, cons, dt, s)
        st.step()
        assert abs(st.dt.item() - dt) <= 1e-12 * dt
    assert H.rel_linf(host(st.primitives)[:, m], prims[:, m]) <= 1e-12


@pytest.mark.parametrize("cells,bc,integrator", [((200, 1, 1), "ZEROGRADIENT", "RK3"), ((48, 40, 1), "PERIODIC", "RK2"),
                                                 ((20, 16, 36), "SYMMETRY", "RK3")])
def test_cuda_graph_step_is_bit_identical(cells, bc, integrator):
    """BlockRuntime.use_cuda_graph: the replayed step (one graph per ping-pong parity) gives the same bits as the
    launched step, including dt / t carried on the device."""
    from jaxfluids_b200 import InputManager, InitializationManager, SimulationManager
    s = H.make_setup(cells, bc=bc, integrator=integrator)
    case = {"general": {"case_name": "g", "end_step": 10, "save_path": "./r"},
            "domain": {ax: {"cells": cells[i], "range": [0.0, 1.0]} for i, ax in enumerate("xyz")},
            "boundary_conditions": {f: {"type": s.bc[f]} for f in port.FACES},
            "initial_condition": {"rho": 1.0, "u": 0.0, "v": 0.0, "w": 0.0, "p": 1.0},
            "material_properties": {"equation_of_state": {"model": "IdealGas", "specific_heat_ratio": 1.4,
                                                          "specific_gas_constant": 1.0}}}
    num = {"conservatives": {"halo_cells": 5, "time_integration": {"integrator": integrator, "CFL": 0.5},
           "convective_fluxes": {"convective_solver": "GODUNOV", "godunov": {"riemann_solver": "HLLC", "signal_speed": "EINFELDT",
           "reconstruction_stencil": "WENO5-Z", "reconstruction_variable": "CHAR-PRIMITIVE"}}},
           "active_physics": {"is_convective_flux": True}, "output": {"logging": {"level": "NONE"}}}
    user = H.smooth_ic(s, seed=17, amp=0.1)[[0] + [1 + i for i in s.active] + [4]]
    outs = []
    for graph in (False, True):
        im = InputManager(case, num)
        buf = InitializationManager(im).initialization(user_prime_init=user)
        rt = SimulationManager(im).runtime
        tcv = buf.time_control_variables
        rt.set_time_control(tcv.physical_simulation_time, tcv.physical_timestep_size)
        if graph:
            rt.use_cuda_graph(True)
        for _ in range(5):
            rt.step()
        outs.append((host(rt.primitives).copy(), host(rt.conservatives).copy(), rt.read_step_scalars()))
    m = H.face_halo_mask(s)
    assert np.array_equal(outs[0][0][:, m], outs[1][0][:, m])
    assert np.array_equal(outs[0][1][:, m], outs[1][1][:, m])
    assert outs[0][2] == outs[1][2]


@pytest.mark.parametrize("cells,gravity,visc", [((16, 40, 1), (0.0, 1.0, 0.0), False), ((12, 14, 24), (0.3, -0.2, 1.0), True),
                                                ((80, 1, 1), (-0.7, 0.0, 0.0), False)])
def test_dirichlet_and_gravity_steps(cells, gravity, visc):
    """DIRICHLET faces with constant primitives (halos/outer/material.py:732-798) on the last active axis, the other
    faces SYMMETRY, plus the gravity source (source_term_solver.py:163-186): rhs incl. volume forces, halo fill
    bit-exact, 4 steps against the oracle (Rayleigh-Taylor-like setups)."""
    from jaxfluids_b200.engine import BlockState
    act = [i for i in range(3) if cells[i] > 1]
    last = act[-1]
    bc = {f: ("DIRICHLET" if port.FACE_AXIS[f] == last else "SYMMETRY") for f in port.FACES}
    s = H.make_setup(cells, bc=bc, recon="PRIMITIVE")
    hi_f, lo_f = port.FACES[2 * last], port.FACES[2 * last + 1]
    s.dirichlet = {hi_f: (1.0, 0.0, 0.0, 0.0, 2.5), lo_f: (2.0, 0.01, -0.02, 0.0, 1.0)}
    s.is_volume_force, s.gravity = True, tuple(gravity)
    s.is_interpolation_limiter = True
    if visc:
        s.is_viscous_flux, s.is_heat_flux, s.dynamic_viscosity = True, True, 5e-3
        s.thermal_conductivity_model, s.prandtl_number = "PRANDTL", 0.72
    prims, cons = port.initialize(H.smooth_ic(s, seed=23, amp=0.05), s)
    sol = make_solver(s)
    p, c = dev(np.nan_to_num(prims, nan=1.0)), dev(np.nan_to_num(cons, nan=1.0))
    scales = H.rhs_scales(prims, s)
    got = host(sol.compute_rhs(p))
    assert H.rel_linf(got, port.compute_rhs(prims, s, cons), scale=scales) <= H.TOL_RHS
    st = BlockState(sol, np.nan_to_num(prims, nan=1.0), np.nan_to_num(cons, nan=1.0))
    dt = port.time_step_size(prims, s)
    m = H.defined_mask(s)
    for _ in range(4):
        prims, cons, dt = port.step(prims, cons, dt, s)
        st.step()
        assert abs(st.dt.item() - dt) <= 1e-12 * dt
    assert H.rel_linf(host(st.primitives)[:, m], prims[:, m]) <= 1e-12
    # the DIRICHLET halos hold exactly the prescribed constants
    gp = host(st.primitives)
    sl = [slice(None)] + list(s.interior)
    sl[1 + last] = slice(0, s.nh)
    assert np.array_equal(gp[tuple(sl)], np.broadcast_to(np.array(s.dirichlet[lo_f]).reshape(5, 1, 1, 1), gp[tuple(sl)].shape))


@pytest.mark.parametrize("cells,bc", [((60, 1, 1), "ZEROGRADIENT"), ((20, 24, 1), "PERIODIC"), ((12, 10, 16), "SYMMETRY")])
def test_dissipative_only_steps(cells, bc):
    """active_physics/is_convective_flux = false (the heat-equation examples): rhs = dissipative fluxes only, stage
    runs unfused (update kernel + halo kernels); 4 steps incl. dt against the oracle."""
    from jaxfluids_b200.engine import BlockState
    s = diss_setup(cells, bc, DISS[1])
    s.is_convective_flux = False
    prims, cons = port.initialize(H.smooth_ic(s, seed=31, amp=0.1), s)
    sol = make_solver(s)
    p = dev(np.nan_to_num(prims, nan=1.0))
    ref = port.compute_rhs(prims, s)
    got = host(sol.compute_rhs(p))
    sc = np.maximum(np.abs(ref).reshape(5, -1).max(axis=1), 1e-3 * np.abs(ref).max())
    assert max(np.abs(got[v] - ref[v]).max() / sc[v] for v in range(5)) <= 1e-12
    st = BlockState(sol, np.nan_to_num(prims, nan=1.0), np.nan_to_num(cons, nan=1.0))
    dt = port.time_step_size(prims, s)
    m = H.defined_mask(s)
    for _ in range(4):
        prims, cons, dt = port.step(prims, cons, dt, s)
        st.step()
        assert abs(st.dt.item() - dt) <= 1e-12 * dt
    assert H.rel_linf(host(st.primitives)[:, m], prims[:, m]) <= 1e-12
    assert H.rel_linf(host(st.conservatives)[:, m], cons[:, m]) <= 1e-12


@pytest.mark.parametrize("sig", ["ARITHMETIC", "RUSANOV", "DAVIS", "TORO"])
@pytest.mark.parametrize("cells,bc,recon", [((120, 1, 1), "ZEROGRADIENT", "CHAR-PRIMITIVE"), ((32, 36, 1), "PERIODIC", "PRIMITIVE"),
                                            ((16, 20, 40), "SYMMETRY", "CHAR-PRIMITIVE")])
def test_hllc_signal_speed_estimates(cells, bc, recon, sig):
    """godunov.signal_speed (signal_speeds.py:10-69, :135-157): per-axis rhs and 3 steps against the oracle."""
    from jaxfluids_b200.engine import BlockState
    s = H.make_setup(cells, bc=bc, recon=recon)
    s.signal_speed = sig
    prims, cons = port.initialize(H.smooth_ic(s, seed=6, amp=0.2), s)
    sol = make_solver(s)
    p = dev(np.nan_to_num(prims, nan=1.0))
    scales = H.rhs_scales(prims, s)
    for a in s.active:
        rhs = sol.new_rhs()
        sol.sweep(a, p, rhs, accumulate=False)
        assert H.rel_linf(host(rhs), port.rhs_axis(prims, a, s), scale=scales) <= H.TOL_RHS, f"axis {a}"
    st = BlockState(sol, np.nan_to_num(prims, nan=1.0), np.nan_to_num(cons, nan=1.0))
    dt = port.time_step_size(prims, s)
    for _ in range(3):
        prims, cons, dt = port.step(prims, cons, dt, s)
        st.step()
    m = H.defined_mask(s)
    assert H.rel_linf(host(st.primitives)[:, m], prims[:, m]) <= 1e-12


@pytest.mark.parametrize("sig", ["EINFELDT", "TORO"])
@pytest.mark.parametrize("cells,bc,recon", [((120, 1, 1), "ZEROGRADIENT", "CHAR-PRIMITIVE"), ((32, 36, 1), "PERIODIC", "PRIMITIVE"),
                                            ((16, 20, 40), "SYMMETRY", "CHAR-PRIMITIVE")])
def test_hll_riemann_solver(cells, bc, recon, sig):
    """riemann_solver = HLL (solvers/riemann_solvers/HLL.py): per-axis rhs and 3 steps against the oracle."""
    from jaxfluids_b200.engine import BlockState
    s = H.make_setup(cells, bc=bc, recon=recon, riemann="HLL")
    s.signal_speed = sig
    prims, cons = port.initialize(H.smooth_ic(s, seed=6, amp=0.2), s)
    sol = make_solver(s)
    p = dev(np.nan_to_num(prims, nan=1.0))
    scales = H.rhs_scales(prims, s)
    for a in s.active:
        rhs = sol.new_rhs()
        sol.sweep(a, p, rhs, accumulate=False)
        assert H.rel_linf(host(rhs), port.rhs_axis(prims, a, s), scale=scales) <= H.TOL_RHS, f"axis {a}"
    st = BlockState(sol, np.nan_to_num(prims, nan=1.0), np.nan_to_num(cons, nan=1.0))
    dt = port.time_step_size(prims, s)
    for _ in range(3):
        prims, cons, dt = port.step(prims, cons, dt, s)
        st.step()
    m = H.defined_mask(s)
    assert H.rel_linf(host(st.primitives)[:, m], prims[:, m]) <= 1e-12


@pytest.mark.parametrize("tag", ["simple", "nasa", "simple_cellsize", "nasa_interp"])
def test_flux_limiter_fixture_rhs(tag):
    """positivity/flux_limiter SIMPLE | NASA (limiter_flux.py:146-330) on the reference's fixture where hundreds of
    faces fall back to the first-order flux (a face switched differently would be an O(1) error): per-axis sweeps,
    compute_rhs, and the public SpaceSolver.compute_rhs with physical_timestep_size."""
    import copy, json
    g = np.load(os.path.join(H.GOLDEN, "special", "flux_limiter_riemann2d_20x24.npz"))
    s = H.setup_from_json(json.loads(str(g[f"case_json_{tag}"])), json.loads(str(g[f"num_json_{tag}"])))
    prims, cons, dt = g[f"prims_halo_{tag}"], g[f"cons_halo_{tag}"], float(g[f"dt_{tag}"])
    sol = make_solver(s)
    p = dev(np.nan_to_num(prims, nan=1.0))
    scales = H.rhs_scales(prims, s)
    with pytest.raises(Exception):                      # no time step bound yet: must fail loudly, not guess
        sol.sweep(s.active[0], p, sol.new_rhs(), accumulate=False)
    sol.bind_timestep(dev(np.array([dt])))
    with np.errstate(all="ignore"):
        for a in s.active:
            rhs = sol.new_rhs()
            sol.sweep(a, p, rhs, accumulate=False)
            assert H.rel_linf(host(rhs), port.rhs_axis(prims, a, s, cons, dt), scale=scales) <= H.TOL_RHS, f"axis {a}"
    got = host(sol.compute_rhs(p))
    assert H.rel_linf(got, g[f"rhs_{tag}"], scale=scales) <= H.TOL_RHS
    # and the limiter matters on this state
    s0 = copy.copy(s)
    s0.flux_limiter = None
    assert H.rel_linf(host(make_solver(s0).compute_rhs(p)), g[f"rhs_{tag}"], scale=scales) > 1e-3
    # a different time step switches a different set of faces
    sol.bind_timestep(dev(np.array([0.5 * dt])))
    with np.errstate(all="ignore"):
        ref_half = port.compute_rhs(prims, s, cons, 0.5 * dt)
    assert H.rel_linf(host(sol.compute_rhs(p)), ref_half, scale=scales) <= H.TOL_RHS


@pytest.mark.parametrize("force_rows", [False, True])
def test_flux_limiter_3d_all_kernels(force_rows, monkeypatch):
    """The limiter through every sweep kernel (march x / y, rows or contig z, fused epilogue) on a 3-D near-vacuum
    state: one stage rhs and 2 steps against the oracle."""
    from jaxfluids_b200.engine import BlockState
    if force_rows:
        monkeypatch.setenv("JXF_FORCE_ROWS", "1")
    s = H.make_setup((20, 18, 40), bc="PERIODIC", recon="CHAR-PRIMITIVE")
    s.flux_limiter = "SIMPLE"
    s.is_interpolation_limiter = True          # keeps the reconstructed states of this state admissible
    ic = H.smooth_ic(s, seed=9, amp=0.3)
    x, y, z = s.cell_centers()
    X, Y, Z = np.meshgrid(x, y, z, indexing="ij")
    ic[0] = np.where(np.sin(2 * np.pi * (X + Y)) * np.cos(2 * np.pi * (Z - Y)) > 0.3, 2e-2, ic[0])   # rarefied pockets
    ic[1:4] *= 4.0                             # (host-simulated: 45 / 265 / 2141 faces switch on x / y / z)
    prims, cons = port.initialize(ic, s)
    dt = 3.0 * port.time_step_size(prims, s)
    sol = make_solver(s)
    sol.bind_timestep(dev(np.array([dt])))
    p = dev(prims)
    scales = H.rhs_scales(prims, s)
    import copy
    s0 = copy.copy(s)
    s0.flux_limiter = None
    with np.errstate(all="ignore"):
        for a in s.active:
            ref = port.rhs_axis(prims, a, s, cons, dt)
            assert np.abs(ref - port.rhs_axis(prims, a, s0)).max() > 0, "limiter inactive on this axis"
            rhs = sol.new_rhs()
            sol.sweep(a, p, rhs, accumulate=False)
            assert H.rel_linf(host(rhs), ref, scale=scales) <= H.TOL_RHS, f"axis {a}"
    # fused stages: one step at this (over-CFL) dt, checked where the oracle's result is finite
    st = BlockState(sol, prims, cons, dt=dt)
    with np.errstate(all="ignore"):
        rp, rc, _ = port.step(prims, cons, dt, s)
    st.step()
    m = H.defined_mask(s) & np.isfinite(rp).all(axis=0)
    assert m.sum() > 0.5 * np.prod(s.cells)
    assert H.rel_linf(host(st.primitives)[:, m], rp[:, m]) <= 1e-11
