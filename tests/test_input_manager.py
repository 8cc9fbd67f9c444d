"""CPU: the reference's shipped case / numerical-setup JSON files select this path unchanged, and
options outside the path fail loudly."""
import copy
import json
import os

import numpy as np
import pytest

from jaxfluids_b200.input_manager import InputManager
from jaxfluids_b200.domain_information import DomainInformation
from tests import helpers as H


def fixture_setups():
    out = {}
    for name in H.golden_names():
        _, case, num = H.load_golden(name)
        out[name] = (case, num)
    return out


SETUPS = fixture_setups()


@pytest.mark.parametrize("name", sorted(SETUPS))
def test_reference_json_parses(name, tmp_path):
    case, num = SETUPS[name]
    im = InputManager(case, num)
    # also through files, like examples/*/run.py
    pc, pn = tmp_path / "case.json", tmp_path / "num.json"
    pc.write_text(json.dumps(case)); pn.write_text(json.dumps(num))
    im2 = InputManager(str(pc), str(pn))
    s = H.setup_from_json(case, num)
    for m in (im, im2):
        di = m.domain_information
        assert di.global_number_of_cells == s.cells
        assert di.device_shape_with_halos == s.shape
        assert tuple(float(x) for x in di.one_cell_sizes) == tuple(float(x) for x in s.inv_dx)
        assert di.smallest_cell_size == float(s.dx_min)
        g = m.numerical_setup.conservatives.convective_fluxes.godunov
        assert (g.reconstruction_variable, g.riemann_solver, g.reconstruction_stencil) == (s.recon, s.riemann, s.stencil)
        assert m.numerical_setup.conservatives.time_integration.integrator == s.integrator
        assert m.case_setup.boundary_condition_setup == s.bc


@pytest.mark.parametrize("name", sorted(SETUPS))
def test_initial_condition_lambdas_match_reference(name):
    g, case, num = H.load_golden(name)
    im = InputManager(case, num)
    mesh = im.domain_information.compute_device_mesh_grid(0)
    ic = im.case_setup.initial_condition_setup
    prims = np.stack([ic[k](*mesh) for k in ("rho", "u", "v", "w", "p")])
    assert prims.shape == g["prims0"].shape
    np.testing.assert_allclose(prims, g["prims0"], rtol=0, atol=1e-15)


def _mod(d, path, value):
    d = copy.deepcopy(d)
    cur = d
    for k in path[:-1]:
        cur = cur.setdefault(k, {})
    cur[path[-1]] = value
    return d


GOD = ("conservatives", "convective_fluxes", "godunov")


@pytest.mark.parametrize("path,value", [
    (GOD + ("riemann_solver",), "LAX-FRIEDRICHS"),
    (GOD + ("riemann_solver",), "CATUM"),
    (GOD + ("signal_speed",), "DAVIS2"),
    (GOD + ("reconstruction_stencil",), "TENO8"),
    (GOD + ("reconstruction_stencil",), "WENO7-JS"),
    (("conservatives", "convective_fluxes", "convective_solver"), "ALDM"),
    (("active_physics", "is_geometric_source"), True),
    (("conservatives", "positivity", "flux_limiter"), "HAS"),
    (("conservatives", "positivity", "flux_partition"), "WAVESPEED"),
    (("conservatives", "positivity", "is_thinc_interpolation_limiter"), True),
    (("precision", "is_double_precision_compute"), False),
])
def test_valid_reference_options_outside_the_path_raise_not_implemented(path, value):
    case, num = SETUPS["tgv16_sym_char_hllc_rk3"]
    with pytest.raises(NotImplementedError, match="B200 path"):
        InputManager(case, _mod(num, path, value))


@pytest.mark.parametrize("stencil", ["WENO1", "WENO3-JS", "WENO3-Z", "WENO3-N", "CENTRAL2", "TENO5", "TENO5-A", "TENO6", "TENO6-A", "WENO6-CU",
                                     "KOREN", "MC", "MINMOD", "SUPERBEE", "VANALBADA", "VANLEER"])
def test_generic_stencil_names_and_rk2_ls4_select_the_path(stencil):
    """The reference's stencil / integrator names select the B200 path unchanged; a halo count that is valid for the
    stencil in the reference but below what the sweep kernels stage says so."""
    case, num = SETUPS["tgv16_sym_char_hllc_rk3"]
    num = _mod(_mod(num, GOD + ("reconstruction_stencil",), stencil), ("conservatives", "time_integration", "integrator"),
               "RK2_LS4")
    im = InputManager(case, num)
    assert im.numerical_setup.conservatives.convective_fluxes.godunov.reconstruction_stencil == stencil
    assert im.numerical_setup.conservatives.time_integration.integrator == "RK2_LS4"
    from jaxfluids_b200 import _lib, registries as R
    assert stencil in _lib.STENCIL and R.REQUIRED_HALOS[stencil] <= 3
    if R.REQUIRED_HALOS[stencil] < 3:
        with pytest.raises(NotImplementedError, match="halo_cells >= 3"):
            InputManager(case, _mod(num, ("conservatives", "halo_cells"), R.REQUIRED_HALOS[stencil]))


@pytest.mark.parametrize("path,value", [
    (GOD + ("riemann_solver",), "HLLD"),
    (GOD + ("reconstruction_stencil",), "WENO5"),
    (("conservatives", "time_integration", "CFL"), -0.5),
    (("conservatives", "halo_cells"), 2),
])
def test_invalid_values_fail_the_reference_consistency_assertion(path, value):
    case, num = SETUPS["tgv16_sym_char_hllc_rk3"]
    with pytest.raises(AssertionError, match="Consistency error in numerical setup file"):
        InputManager(case, _mod(num, path, value))


def test_dissipative_setup_is_read_like_the_reference():
    """active_physics + dissipative_fluxes + material_properties/transport (read_conservatives.py:374-440,
    read_material_manager.py:200-330)."""
    case, num = SETUPS["tgv12_sym_visc_prandtl_rk3"]
    im = InputManager(case, num)
    ap = im.numerical_setup.active_physics
    assert ap.is_viscous_flux and ap.is_heat_flux and ap.is_viscous_heat_production
    df = im.numerical_setup.conservatives.dissipative_fluxes
    assert (df.reconstruction_stencil, df.derivative_stencil_center, df.derivative_stencil_face) == ("CENTRAL4",) * 3
    tr = im.case_setup.material_setup.transport
    assert tr.dynamic_viscosity == 1 / 160 and tr.thermal_conductivity_model == "PRANDTL" and tr.prandtl_number == 0.71
    # valid reference options this path does not implement
    with pytest.raises(NotImplementedError, match="B200 path"):
        InputManager(case, _mod(num, ("conservatives", "dissipative_fluxes", "reconstruction_stencil"), "CENTRAL6"))
    with pytest.raises(NotImplementedError, match="B200 path"):
        InputManager(case, _mod(num, ("conservatives", "dissipative_fluxes", "is_laplacian"), True))
    with pytest.raises(NotImplementedError, match="B200 path"):
        InputManager(_mod(case, ("material_properties", "transport", "dynamic_viscosity"),
                          {"model": "SUTHERLAND", "sutherland_parameters": [1.7e-5, 273.0, 110.4]}), num)
    # missing transport block with the viscous flux on: the reference's "not optional" assertion
    bad = copy.deepcopy(case)
    del bad["material_properties"]["transport"]
    with pytest.raises(AssertionError, match="Consistency error in case setup file"):
        InputManager(bad, num)


def test_case_errors():
    case, num = SETUPS["tgv16_sym_char_hllc_rk3"]
    with pytest.raises(NotImplementedError):
        InputManager(_mod(case, ("boundary_conditions", "east", "type"), "LINEAREXTRAPOLATION"), num)
    # DIRICHLET: primitives_callable
    with pytest.raises(AssertionError, match="primitives_callable"):
        InputManager(_mod(case, ("boundary_conditions", "east", "type"), "DIRICHLET"), num)
    dirich = _mod(case, ("boundary_conditions", "east"),
                  {"type": "DIRICHLET", "primitives_callable": {"rho": 1.0, "u": 0.0, "v": 0.0, "w": 0.0, "p": 2.5}})
    assert InputManager(dirich, num).case_setup.dirichlet_setup == {"east": (1.0, 0.0, 0.0, 0.0, 2.5)}
    # gravity needs forcings/gravity when is_volume_force is on
    with pytest.raises(AssertionError, match="gravity"):
        InputManager(case, _mod(num, ("active_physics", "is_volume_force"), True))
    grav = _mod(case, ("forcings",), {"gravity": [0.0, 1.0, 0.0]})
    assert InputManager(grav, _mod(num, ("active_physics", "is_volume_force"), True)).case_setup.gravity == (0.0, 1.0, 0.0)
    # WALL: wall_velocity_callable is read (floats or lambda strings); a missing one is the reference's consistency error
    with pytest.raises(AssertionError, match="wall_velocity_callable"):
        InputManager(_mod(case, ("boundary_conditions", "east", "type"), "WALL"), num)
    wall = _mod(case, ("boundary_conditions", "east"), {"type": "WALL", "wall_velocity_callable": {"u": 0.0, "v": 0.5, "w": 0.0}})
    assert InputManager(wall, num).case_setup.wall_velocity_setup == {"east": (0.0, 0.5, 0.0)}
    lam = InputManager(_mod(wall, ("boundary_conditions", "east", "wall_velocity_callable", "v"), "lambda y, z, t: 0.5"), num)
    assert lam.case_setup.wall_velocity_setup["east"] == (0.0, "lambda y, z, t: 0.5", 0.0)   # evaluated by the runtime
    with pytest.raises(AssertionError, match="case setup"):
        InputManager(_mod(case, ("boundary_conditions", "east", "type"), "PERIODIC"), num)   # west is SYMMETRY
    with pytest.raises(AssertionError, match="argument labels"):
        InputManager(_mod(case, ("initial_condition", "u"), "lambda x, y: x"), num)
    with pytest.raises(NotImplementedError):
        InputManager(_mod(case, ("material_properties", "equation_of_state", "model"), "StiffenedGas"), num)
    bad = copy.deepcopy(case); del bad["general"]["end_time"]
    with pytest.raises(AssertionError, match="end_time or end_step"):
        InputManager(bad, num)


def test_decomposition_bookkeeping():
    di = DomainInformation((64, 32, 16), ((0, 1),) * 3, (2, 2, 2), 5)
    assert di.device_number_of_cells == (32, 16, 8)
    # rank = i*sy*sz + j*sz + k  (domain/helper_functions.py:155-169)
    assert [di.block_index(r) for r in range(8)] == [(i, j, k) for i in range(2) for j in range(2) for k in range(2)]
    assert di.neighbor(0, "east", periodic=False) == 4 and di.neighbor(0, "west", periodic=False) is None
    assert di.neighbor(0, "west", periodic=True) == 4
    assert di.neighbor(5, "north", periodic=False) == 7 and di.neighbor(5, "top", periodic=False) is None
    assert di.neighbor(5, "bottom", periodic=False) == 4
    assert di.block_slices(6) == (slice(32, 64), slice(16, 32), slice(0, 8))
    di1 = DomainInformation((64, 1, 1), ((0, 1),) * 3, (1, 1, 1), 5)
    assert di1.neighbor(0, "east", periodic=True) is None     # unsplit periodic axis is a local BC


def test_flux_splitting_block_is_read_like_the_reference():
    """convective_solver = FLUX-SPLITTING (read_conservatives.py:205-244): the flux_splitting block selects the path, the
    godunov block is not needed; CLF (accepted by the reference's input check, unhandled by its eigendecomposition)
    says 'not implemented'."""
    case, num = SETUPS["tgv16_sym_char_hllc_rk3"]
    num = copy.deepcopy(num)
    cf = num["conservatives"]["convective_fluxes"]
    cf["convective_solver"] = "FLUX-SPLITTING"
    cf.pop("godunov")
    cf["flux_splitting"] = {"flux_splitting": "CLLF", "reconstruction_stencil": "WENO6-CU"}
    im = InputManager(case, num)
    c = im.numerical_setup.conservatives.convective_fluxes
    assert c.convective_solver == "FLUX-SPLITTING" and c.godunov is None
    assert (c.flux_splitting.flux_splitting, c.flux_splitting.reconstruction_stencil, c.flux_splitting.frozen_state) == \
        ("CLLF", "WENO6-CU", "ARITHMETIC")
    roe = InputManager(case, _mod(num, ("conservatives", "convective_fluxes", "flux_splitting", "frozen_state"), "ROE"))
    assert roe.numerical_setup.conservatives.convective_fluxes.flux_splitting.frozen_state == "ROE"
    for key, value in (("flux_splitting", "CLF"), ("reconstruction_stencil", "WENO7-JS")):
        with pytest.raises(NotImplementedError, match="B200 path"):
            InputManager(case, _mod(num, ("conservatives", "convective_fluxes", "flux_splitting", key), value))
    with pytest.raises(AssertionError, match="Consistency error in numerical setup file"):
        InputManager(case, _mod(num, ("conservatives", "convective_fluxes", "flux_splitting", "flux_splitting"), "HLL"))
    with pytest.raises(NotImplementedError, match="FLUX-SPLITTING"):
        InputManager(case, _mod(num, ("conservatives", "positivity", "flux_limiter"), "SIMPLE"))


@pytest.mark.parametrize("name,expect", [
    ("generic/lax100_fs_roe_weno6cu_rk3", dict(convective_solver=1, flux_splitting=1, stencil=6, integrator=2)),
    ("generic/riemann2d_16x20_fs_cllf_teno5_rk3", dict(convective_solver=1, flux_splitting=2, stencil=5)),
    ("generic/sod100_teno5_char_hllc_rk2ls4", dict(convective_solver=0, stencil=5, recon=1, riemann=0, integrator=3)),
    ("generic/riemann2d_20x24_prim_ausmp_rk3", dict(convective_solver=0, stencil=0, recon=0, riemann=4)),
    ("generic/sod100_char_hllclm_rk3", dict(riemann=3, signal_speed=0)),
    ("generic/sod100_vanleer_prim_rusanov_rk3", dict(stencil=12, recon=0, riemann=1)),
    ("sod200_char_hllc_rk3", dict(convective_solver=0, stencil=0, recon=1, riemann=0, integrator=2, frozen_state=0)),
    ("generic/sod100_charcons_roe_hllc_rk3", dict(recon=3, frozen_state=1, stencil=0)),
    ("generic/riemann2d_16x20_cons_teno5_hll_rk3", dict(recon=2, frozen_state=0, stencil=5, riemann=2)),
    ("generic/lax100_fs_roe_weno6cu_roefrozen_rk3", dict(convective_solver=1, flux_splitting=1, stencil=6, frozen_state=1)),
    ("api/heat2d_24x20_dirichlet_lambda_noconv_rk3", dict(no_convective_flux=1, heat_flux=1, stencil=1)),
    ("api/riemann2d_16x20_inflow_outflow_visc_rk3", dict(viscous_flux=1, heat_flux=1)),
    ("api/dmr_48x32_dirichlet_symmetry_south_rk3", dict(interpolation_limiter=1, recon=1)),
    ("api/cavity_24x20_wall_lambda_lid_visc_rk3", dict(viscous_flux=1, stencil=1)),
])

def test_json_options_reach_the_c_config(name, expect, monkeypatch):
    """JSON -> InputManager -> BlockRuntime -> BlockConfig.to_c(): the ids the C ABI receives (include/jxf_b200.h),
    checked without a GPU by stopping at the solver's construction."""
    import jaxfluids_b200.runtime as RT
    from jaxfluids_b200.parallel import ParallelContext

    class Reached(Exception):
        pass

    def fake_solver(cfg, *a, **k):
        raise Reached(cfg.to_c())
    monkeypatch.setattr(RT, "BlockSolver", fake_solver)
    g, case, num = H.load_golden(name)
    im = InputManager(case, num)
    with pytest.raises(Reached) as info:
        RT.BlockRuntime(im, ParallelContext(im.domain_information))
    c = info.value.args[0]
    for key, value in expect.items():
        assert getattr(c, key) == value, key
    # NEUMANN / SIMPLE_INFLOW / SIMPLE_OUTFLOW faces reach the kernels as ZEROGRADIENT (3); the host applies their data
    for k, f in enumerate(("east", "west", "north", "south", "top", "bottom")):
        entry = case["boundary_conditions"][f]
        if isinstance(entry, list):                          # several types: the kernels fill the SYMMETRY entry (2)
            assert c.bc[k] == 2
        elif entry["type"] in ("NEUMANN", "SIMPLE_INFLOW", "SIMPLE_OUTFLOW"):
            assert c.bc[k] == 3


@pytest.mark.parametrize("variable", ["PRIMITIVE", "CONSERVATIVE", "CHAR-PRIMITIVE", "CHAR-CONSERVATIVE"])
@pytest.mark.parametrize("frozen", ["ARITHMETIC", "ROE"])
def test_every_reconstruction_variable_and_frozen_state_selects_the_path(variable, frozen):
    case, num = SETUPS["tgv16_sym_char_hllc_rk3"]
    num = _mod(_mod(num, GOD + ("reconstruction_variable",), variable), GOD + ("frozen_state",), frozen)
    g = InputManager(case, num).numerical_setup.conservatives.convective_fluxes.godunov
    assert (g.reconstruction_variable, g.frozen_state) == (variable, frozen)


@pytest.mark.parametrize("name", H.golden_names() + H.generic_golden_names() + H.api_golden_names())
def test_block_config_of_every_fixture_matches_the_oracle_setup(name, monkeypatch):
    """JSON -> InputManager -> BlockRuntime -> BlockConfig, field by field against the oracle's reading of the same JSON
    (tests/helpers.setup_from_json), for every fixture: what the kernels are configured with is what the oracle runs."""
    import jaxfluids_b200.runtime as RT
    from jaxfluids_b200.parallel import ParallelContext

    class Reached(Exception):
        pass

    def fake_solver(cfg, *a, **k):
        raise Reached(cfg)
    monkeypatch.setattr(RT, "BlockSolver", fake_solver)
    g, case, num = H.load_golden(name)
    im = InputManager(case, num)
    with pytest.raises(Reached) as info:
        RT.BlockRuntime(im, ParallelContext(im.domain_information))
    cfg, s = info.value.args[0], H.setup_from_json(case, num)
    assert tuple(cfg.cells) == tuple(s.cells) and cfg.nh == s.nh and cfg.gamma == s.gamma
    assert tuple(np.float64(x) for x in cfg.inv_dx) == tuple(s.inv_dx) and np.float64(cfg.dx_min) == s.dx_min
    assert (cfg.convective_solver, cfg.stencil, cfg.integrator, cfg.cfl) == (s.convective_solver, s.stencil, s.integrator, s.cfl)
    assert cfg.frozen_state == s.frozen_state
    if s.convective_solver == "FLUX-SPLITTING":
        assert cfg.flux_splitting == s.flux_splitting
    else:
        assert (cfg.recon, cfg.riemann, cfg.signal_speed) == (s.recon, s.riemann, s.signal_speed)
    assert (cfg.is_interpolation_limiter, cfg.limit_velocity, cfg.flux_limiter, cfg.flux_partition) == \
        (s.is_interpolation_limiter, s.limit_velocity, s.flux_limiter, s.flux_partition)
    assert (cfg.is_viscous_flux, cfg.is_heat_flux, cfg.is_convective_flux, cfg.is_volume_force) == \
        (s.is_viscous_flux, s.is_heat_flux, s.is_convective_flux, s.is_volume_force)
    if s.is_viscous_flux:                                     # (the viscosity is not read with the heat flux alone)
        assert (cfg.dynamic_viscosity, cfg.bulk_viscosity) == (s.dynamic_viscosity, s.bulk_viscosity)
    if s.is_dissipative:
        assert cfg.gas_constant == s.gas_constant
    assert tuple(cfg.gravity) == tuple(s.gravity)
    kernel_type = {"NEUMANN": "ZEROGRADIENT", "SIMPLE_INFLOW": "ZEROGRADIENT", "SIMPLE_OUTFLOW": "ZEROGRADIENT"}
    assert {f: cfg.bc[f] for f in H.port.FACES} == {f: kernel_type.get(s.bc[f], s.bc[f]) for f in H.port.FACES}
    for f, vals in s.dirichlet.items():                      # constants reach the kernels; arrays are host-applied
        if all(isinstance(v, float) for v in vals):
            assert tuple(cfg.dirichlet[f]) == tuple(vals)
    for f, vals in s.wall_velocity.items():
        if all(isinstance(v, float) for v in vals):
            assert tuple(cfg.wall_velocity[f]) == tuple(vals)
