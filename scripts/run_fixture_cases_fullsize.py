"""Sanity run (not a parity test): the case files stored in the fixtures (the reference's shipped examples, shrunk
for the fixtures) at their ORIGINAL grid sizes for a few hundred steps through the public API; prints step count,
time, dt, min rho / min p and MCUPS.  Usage: python scripts/run_fixture_cases_fullsize.py [steps]"""
import copy, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from tests import helpers as H
from jaxfluids_b200 import InputManager, InitializationManager, SimulationManager

FULL = {"cavity_24x20_wall_js_visc_rk3": (128, 128, 1), "rti_16x48_dirichlet_gravity_rk3": (64, 256, 1),
        "heat1d_40_dirichlet_noconv_rk3": (100, 1, 1), "sod200_char_hllc_rk3": (1000, 1, 1),
        "riemann2d_32x32_char_hllc_rk3": (1024, 1024, 1), "tgv12_sym_visc_prandtl_rk3": (128, 128, 128),
        # the examples added with the generic stencils / flux splitting / host-applied boundaries
        "generic/lax100_fs_roe_weno6cu_rk3": (200, 1, 1), "generic/woodward200_fs_roe_weno5z_rk3": (400, 1, 1),
        "api/heat2d_24x20_dirichlet_lambda_noconv_rk3": (100, 100, 1),
        "api/dmr_48x32_dirichlet_symmetry_south_rk3": (256, 256, 1),
        "generic/tgv_10x8x12_per_minmod_char_hllc_rk2ls4": (64, 64, 64), "generic/tgv_10x8x12_sym_char_hllclm_rk3": (64, 64, 64)}
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
for name, cells in FULL.items():
    _, case, num = H.load_golden(name)
    case, num = copy.deepcopy(case), copy.deepcopy(num)
    for ax, n in zip("xyz", cells):
        case["domain"][ax]["cells"] = n
    case["general"]["end_step"] = steps
    case["general"]["end_time"] = 1e9
    num.setdefault("output", {}).setdefault("logging", {})["level"] = "NONE"
    im = InputManager(case, num)
    buf = InitializationManager(im).initialization()
    sim = SimulationManager(im)
    torch.cuda.synchronize()
    t0 = time.time()
    sim.simulate(buf)
    torch.cuda.synchronize()
    el = time.time() - t0
    out = sim.final_buffers
    tcv = out.time_control_variables
    pos = out.step_information.positivity[-1]
    p = out.simulation_buffers.material_fields.primitives
    ok = bool(torch.isfinite(p[(slice(None),) + sim.runtime.cfg.interior]).all())
    print(json.dumps({"case": name, "cells": cells, "steps": tcv.simulation_step, "t": tcv.physical_simulation_time,
                      "dt": tcv.physical_timestep_size, "min_rho": pos.min_density, "min_p": pos.min_pressure,
                      "finite": ok, "MCUPS_incl_host_loop": int(np.prod(cells)) * steps / el / 1e6}), flush=True)
    assert ok and pos.min_density > 0 and pos.min_pressure > 0
