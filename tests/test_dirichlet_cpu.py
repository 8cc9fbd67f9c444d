"""CPU: space-dependent DIRICHLET data (primitives_callable given as lambdas of the transverse coordinates,
halos/outer/material.py:770-790) -- the host-side pieces of the B200 path against the pinned oracle: the case-file
evaluation on the block's transverse cells and the per-face data arrays / op codes BlockRuntime hands to the kernels
(jxf_set_face_data), applied here by a NumPy statement of what the kernels do with them (helpers.apply_face_data_numpy;
the kernels themselves are checked on the GPU against the tests/golden/api fixtures)."""
import copy
import json
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import port
from tests import helpers as H

HEAT2D = {
    "general": {"case_name": "heat2d", "end_time": 1.0, "save_path": "./results", "save_dt": 1.0},
    "domain": {"x": {"cells": 20, "range": [0.0, 1.0]}, "y": {"cells": 16, "range": [0.0, 1.0]},
               "z": {"cells": 1, "range": [0.0, 1.0]},
               "decomposition": {"split_x": 1, "split_y": 1, "split_z": 1}},
    "boundary_conditions": {
        "east": {"type": "DIRICHLET", "primitives_callable": {"rho": 1.0, "u": 0.0, "v": 0.0, "w": 0.0, "p": 1.0}},
        "west": {"type": "DIRICHLET", "primitives_callable": {"rho": "lambda y,t: 1.0 + 0.1 * jnp.cos(3 * y)", "u": 0.0,
                                                              "v": "lambda y,t: 0.2 * y", "w": 0.0, "p": 1.0}},
        "north": {"type": "DIRICHLET", "primitives_callable": {"rho": 1.0, "u": 0.0, "v": 0.0, "w": 0.0,
                                                               "p": "lambda x,t: 1.0 + jnp.sin(jnp.pi * x)"}},
        "south": {"type": "DIRICHLET", "primitives_callable": {"rho": 1.0, "u": 0.0, "v": 0.0, "w": 0.0, "p": 1.0}},
        "top": {"type": "INACTIVE"}, "bottom": {"type": "INACTIVE"}},
    "initial_condition": {"rho": 1.0, "u": 0.0, "v": 0.0, "w": 0.0, "p": 1.0},
    "material_properties": {"equation_of_state": {"model": "IdealGas", "specific_heat_ratio": 1.4,
                                                  "specific_gas_constant": 1.0},
                            "transport": {"dynamic_viscosity": {"model": "CUSTOM", "value": 0.1}, "bulk_viscosity": 0.0,
                                          "thermal_conductivity": {"model": "CUSTOM", "value": 0.1}}},
}
NUM = {"conservatives": {"halo_cells": 4, "time_integration": {"integrator": "RK3", "CFL": 0.9},
                         "convective_fluxes": {"convective_solver": "GODUNOV", "godunov": {
                             "riemann_solver": "HLLC", "signal_speed": "EINFELDT", "reconstruction_stencil": "WENO5-JS",
                             "reconstruction_variable": "PRIMITIVE"}},
                         "dissipative_fluxes": {"reconstruction_stencil": "CENTRAL4", "derivative_stencil_center": "CENTRAL4",
                                                "derivative_stencil_face": "CENTRAL4"}},
       "active_physics": {"is_convective_flux": False, "is_viscous_flux": False, "is_heat_flux": True, "is_volume_force": False},
       "output": {"logging": {"level": "NONE"}}}


def _host_runtime(im, s, **cfg_kw):
    """A BlockRuntime with everything but the CUDA solver: the host-side boundary logic on CPU tensors."""
    from jaxfluids_b200.engine import BlockConfig
    from jaxfluids_b200.parallel import ParallelContext
    from jaxfluids_b200.runtime import BlockRuntime
    rt = BlockRuntime.__new__(BlockRuntime)
    rt.bc_block = dict(im.case_setup.boundary_condition_setup)
    rt._cell_sizes = tuple(im.domain_information.cell_sizes)
    consts = rt._dirichlet_constants(im.case_setup, im.domain_information, ParallelContext(im.domain_information))
    rt.wall_constants = rt._wall_constants(im.case_setup, im.domain_information, ParallelContext(im.domain_information))
    rt.cfg = BlockConfig(cells=s.cells, inv_dx=tuple(float(x) for x in s.inv_dx), dx_min=float(s.dx_min), gamma=s.gamma,
                         bc=rt._kernel_boundary_types(), nh=s.nh, dirichlet=consts, wall_velocity=rt.wall_constants,
                         **cfg_kw)
    rt.cfg.to_c()                                            # the kernels' configuration is constructible
    rt.device = torch.device("cpu")
    edges = []
    rt.solver = SimpleNamespace(active=s.active, halo_fill_edges=lambda p, c: edges.append(1))
    rt.face_data = rt._make_face_data()

    def apply(tp, tc):
        """the kernels' application of the face data on tensors holding the base-rule halos"""
        p, c = H.apply_face_data_numpy(rt.face_data, tp.numpy(), tc.numpy(), s.cells, s.nh, s.gamma)
        tp.copy_(torch.as_tensor(p))
        tc.copy_(torch.as_tensor(c))
    rt._apply_host_boundaries = apply
    return rt, consts, edges


def test_dirichlet_lambdas_are_evaluated_like_the_oracle():
    from jaxfluids_b200.input_manager import InputManager, evaluate_dirichlet_face
    im = InputManager(HEAT2D, NUM)
    s = H.setup_from_json(HEAT2D, NUM)
    for face in ("east", "west", "north", "south"):
        got = evaluate_dirichlet_face(im.case_setup.dirichlet_setup[face], face, im.domain_information)
        for a, b in zip(got, s.dirichlet[face]):
            assert type(a) is type(b) or (isinstance(a, np.ndarray) and isinstance(b, np.ndarray))
            assert np.array_equal(a, b)
    assert isinstance(s.dirichlet["north"][4], np.ndarray) and s.dirichlet["north"][4].shape == (20, 1, 1)
    assert isinstance(s.dirichlet["west"][0], np.ndarray) and s.dirichlet["west"][0].shape == (1, 16, 1)


def test_time_dependent_dirichlet_and_wrong_labels_are_refused():
    from jaxfluids_b200.input_manager import InputManager, evaluate_dirichlet_face
    case = copy.deepcopy(HEAT2D)
    case["boundary_conditions"]["north"]["primitives_callable"]["p"] = "lambda x,t: 1.0 + t * x"
    im = InputManager(case, NUM)
    with pytest.raises(NotImplementedError, match="time-dependent"):
        evaluate_dirichlet_face(im.case_setup.dirichlet_setup["north"], "north", im.domain_information)
    case["boundary_conditions"]["north"]["primitives_callable"]["p"] = "lambda y,t: 1.0 + y"
    im = InputManager(case, NUM)
    with pytest.raises(AssertionError, match="Input argument labels"):
        evaluate_dirichlet_face(im.case_setup.dirichlet_setup["north"], "north", im.domain_information)


def test_halo_slabs_reproduce_the_oracle_halo_fill():
    """BlockRuntime's slab construction and overwrite on CPU tensors: after the halo kernel's placeholder fill (here: the
    oracle's fill with the placeholder constants the kernels get), the face halos of primitives AND conservatives are
    bit-identical to the oracle's fill with the space-dependent data."""
    from jaxfluids_b200.engine import BlockConfig
    from jaxfluids_b200.input_manager import InputManager
    from jaxfluids_b200.parallel import ParallelContext
    from jaxfluids_b200.runtime import BlockRuntime
    im = InputManager(HEAT2D, NUM)
    s = H.setup_from_json(HEAT2D, NUM)
    rt, consts, edges = _host_runtime(im, s, is_heat_flux=True)
    assert set(rt._host_faces) == {"west", "north"} and set(consts) == {"east", "west", "north", "south"}
    assert all(isinstance(v, float) for vals in consts.values() for v in vals)
    # the state a halo kernel leaves: placeholder constants on the varying faces
    rng = np.random.default_rng(3)
    interior = 1.0 + 0.1 * rng.random((5,) + s.cells)
    prims, cons = port.initialize(interior, s)
    s_placeholder = copy.copy(s)
    s_placeholder.dirichlet = consts
    s_placeholder.is_heat_flux = False                       # faces only (the edges follow the slabs on the device)
    p0, c0 = port.halo_fill(prims, cons, s_placeholder)
    s_faces = copy.copy(s)
    s_faces.is_heat_flux = False
    ref_p, ref_c = port.halo_fill(prims, cons, s_faces)
    assert not np.array_equal(p0, ref_p)
    tp, tc = torch.as_tensor(p0.copy()), torch.as_tensor(c0.copy())
    rt._apply_host_boundaries(tp, tc)
    m = H.face_halo_mask(s)
    assert np.array_equal(tp.numpy()[:, m], ref_p[:, m])
    assert np.array_equal(tc.numpy()[:, m], ref_c[:, m])


def test_neumann_and_simple_inflow_outflow_reproduce_the_oracle_halo_fill():
    """NEUMANN (halos/outer/material.py:825-866), SIMPLE_INFLOW (:966-1022), SIMPLE_OUTFLOW (:1024-1050): the kernels are
    configured with ZEROGRADIENT on these faces; the host's index assignments on top of that fill give the oracle's
    halos bit for bit (primitives and conservatives)."""
    from jaxfluids_b200.input_manager import InputManager
    case = copy.deepcopy(HEAT2D)
    case["boundary_conditions"].update({
        "west": {"type": "SIMPLE_INFLOW", "primitives_callable": {"rho": "lambda y,t: 0.5 + 0.2 * y", "u": 1.2,
                                                                  "v": "lambda y,t: 0.1 * jnp.sin(6 * y)", "w": 0.0}},
        "east": {"type": "SIMPLE_OUTFLOW", "primitives_callable": {"p": "lambda y,t: 1.0 + 0.5 * y"}},
        "north": {"type": "NEUMANN", "primitives_callable": {"rho": 0.1, "u": "lambda x,t: 0.2 * x", "v": 0.0, "w": 0.0,
                                                             "p": -0.3}},
        "south": {"type": "NEUMANN", "primitives_callable": {"rho": "lambda x,t: 0.3 * jnp.cos(5 * x)", "u": 0.0, "v": 0.1,
                                                             "w": 0.0, "p": 0.2}}})
    num = copy.deepcopy(NUM)
    num["active_physics"] = {"is_convective_flux": True}
    im = InputManager(case, num)
    s = H.setup_from_json(case, num)
    assert s.bc["west"] == "SIMPLE_INFLOW" and s.bc_values["east"][:4] == (None,) * 4
    rt, consts, edges = _host_runtime(im, s)
    assert consts == {} and set(rt._host_faces) == {"west", "east", "north", "south"}
    assert set(rt.cfg.bc[f] for f in ("west", "east", "north", "south")) == {"ZEROGRADIENT"}
    rng = np.random.default_rng(5)
    prims, cons = port.initialize(1.0 + 0.1 * rng.random((5,) + s.cells), s)
    s_kernel = copy.copy(s)
    s_kernel.bc = dict(rt.cfg.bc)                            # what the halo kernel fills: ZEROGRADIENT
    p0, c0 = port.halo_fill(prims, cons, s_kernel)
    ref_p, ref_c = port.halo_fill(prims, cons, s)
    tp, tc = torch.as_tensor(p0.copy()), torch.as_tensor(c0.copy())
    rt._apply_host_boundaries(tp, tc)
    m = H.face_halo_mask(s)
    assert not np.array_equal(p0[:, m], ref_p[:, m])
    assert np.array_equal(tp.numpy()[:, m], ref_p[:, m])
    assert np.array_equal(tc.numpy()[:, m], ref_c[:, m])


def test_several_types_on_one_face_reproduce_the_oracle_halo_fill():
    """The shipped double Mach reflection example (shrunk): the south face is DIRICHLET for x < 1/6 and SYMMETRY beyond
    (a list with bounding_domain lambdas, halos/outer/material.py:121-277).  The kernels fill SYMMETRY on the whole
    face, the host writes the DIRICHLET state inside its bounding domain: halos bit-identical to the oracle's."""
    from jaxfluids_b200.input_manager import InputManager
    g, case, num = H.load_golden("api/dmr_48x32_dirichlet_symmetry_south_rk3")
    im = InputManager(case, num)
    assert im.case_setup.boundary_condition_setup["south"] == "SYMMETRY" and "south" in im.case_setup.multi_type_setup
    s = H.setup_from_json(case, num)
    assert s.bc["south"] == "SYMMETRY" and [e["kind"] for e in s.bc_multi["south"]] == ["DIRICHLET", "SYMMETRY"]
    rt, consts, edges = _host_runtime(im, s)
    assert set(rt._host_faces) == {("south", 0)} and set(consts) == {"west"}
    prims, cons = g["prims0_halo"], g["cons0_halo"]
    s_kernel = copy.copy(s)
    s_kernel.bc_multi = {}                                   # what the halo kernel fills: SYMMETRY on the whole face
    p0, c0 = port.halo_fill(prims, cons, s_kernel)
    ref_p, ref_c = port.halo_fill(prims, cons, s)
    m = H.face_halo_mask(s)
    assert not np.array_equal(p0[:, m], ref_p[:, m])
    tp, tc = torch.as_tensor(p0.copy()), torch.as_tensor(c0.copy())
    rt._apply_host_boundaries(tp, tc)
    assert np.array_equal(tp.numpy()[:, m], ref_p[:, m])
    assert np.array_equal(tc.numpy()[:, m], ref_c[:, m])
    # bounding domains that do not partition the face are refused
    bad = copy.deepcopy(case)
    bad["boundary_conditions"]["south"][1]["bounding_domain"] = "lambda x: x >= 0.5"
    im_bad = InputManager(bad, num)
    with pytest.raises(NotImplementedError, match="partition"):
        _host_runtime(im_bad, s)


def test_space_dependent_wall_velocity_reproduces_the_oracle_halo_fill():
    """WALL with wall_velocity_callable lambdas (halos/outer/material.py:473-520; a regularised lid): the kernels run with
    wall velocity 0 on such a face (halo = -u_mirror), the host adds 2 u_wall: bit-identical to the oracle's halos."""
    from jaxfluids_b200.input_manager import InputManager
    g, case, num = H.load_golden("cavity_24x20_wall_js_visc_rk3")
    case = copy.deepcopy(case)
    case["boundary_conditions"]["north"]["wall_velocity_callable"] = {"u": "lambda x,t: 16.0 * x**2 * (1.0 - x)**2",
                                                                      "v": 0.0, "w": 0.0}
    case["boundary_conditions"]["east"]["wall_velocity_callable"] = {"u": 0.0, "v": "lambda y,t: 0.1 * jnp.sin(3 * y)",
                                                                     "w": 0.0}
    im = InputManager(case, num)
    s = H.setup_from_json(case, num)
    assert isinstance(s.wall_velocity["north"][0], np.ndarray) and s.wall_velocity["south"] == (0.0, 0.0, 0.0)
    rt, consts, edges = _host_runtime(im, s, is_viscous_flux=True)
    assert set(rt._host_faces) == {"north", "east"} and rt.wall_constants["north"] == (0.0, 0.0, 0.0)
    assert rt.cfg.bc["north"] == "WALL"
    rng = np.random.default_rng(7)
    interior = 1.0 + 0.1 * rng.random((5,) + s.cells)
    interior[1:4] -= 1.05
    s_faces = copy.copy(s)
    s_faces.is_viscous_flux = False                          # faces only (the edges follow on the device)
    prims, cons = port.initialize(interior, s_faces)
    s_kernel = copy.copy(s_faces)
    s_kernel.wall_velocity = dict(rt.wall_constants)         # what the halo kernel fills: u_wall = 0 on north / east
    p0, c0 = port.halo_fill(prims, cons, s_kernel)
    ref_p, ref_c = port.halo_fill(prims, cons, s_faces)
    m = H.face_halo_mask(s)
    assert not np.array_equal(p0[:, m], ref_p[:, m])
    tp, tc = torch.as_tensor(p0.copy()), torch.as_tensor(c0.copy())
    rt._apply_host_boundaries(tp, tc)
    assert np.array_equal(tp.numpy()[:, m], ref_p[:, m])
    assert np.array_equal(tc.numpy()[:, m], ref_c[:, m])
