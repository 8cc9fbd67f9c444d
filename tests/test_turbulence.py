"""Turbulent initial conditions (jaxfluids_b200/turbulence.py) and the HIT-like fixture.

CPU: the HIT generator against the reference's own (turbulence/initialization/hit.py, run on the stand-in when
/root/reference exists), properties of the benchmark's synthetic field, the oracle on the fixture generated from the
reference.  GPU: the kernels on the same fixture.
"""
import json
import os

import numpy as np
import pytest

from oracle import port
from tests import helpers as H
from jaxfluids_b200 import turbulence as T

FIX = os.path.join(H.GOLDEN, "special", "hit32_per_char_hllc_rk3.npz")


def _fixture_setup():
    g = np.load(FIX)
    case, num = json.loads(str(g["case_json"])), json.loads(str(g["num_json"]))
    return g, case, num, H.setup_from_json(case, num)


def test_synthetic_field_is_solenoidal_and_has_the_target_mach_number_and_spectrum():
    n = 48
    p = T.synthetic_solenoidal_ic(n, gamma=1.4, k0=4.0, ma_t=0.4, seed=0)
    assert p.shape == (5, n, n, n) and np.all(p[0] == 1.0) and np.all(p[4] == 1.0 / 1.4)
    v = p[1:4]
    assert abs(np.sqrt(np.mean((v ** 2).sum(0))) - 0.4) <= 1e-12          # q_rms / c, c = 1
    k = np.fft.fftfreq(n, 1 / n)
    kx, ky, kz = np.meshgrid(k, k, k, indexing="ij")
    vh = np.fft.fftn(v, axes=(1, 2, 3))
    assert np.abs(kx * vh[0] + ky * vh[1] + kz * vh[2]).max() <= 1e-12 * np.abs(vh).max()
    spec = T.energy_spectrum_spectral(np.fft.rfftn(v, axes=(1, 2, 3)), n, 0.5)
    assert spec.argmax() == 4 and spec[13:].max() <= 1e-12 * spec.max()   # peak at k0, band-limited to |k| <= 12


def test_synthetic_field_blocks_are_slices_of_the_global_field():
    n = 32
    full = T.synthetic_solenoidal_ic(n)
    for blk in ((slice(0, 16), slice(0, 32), slice(0, 32)), (slice(16, 32), slice(8, 24), slice(0, 16))):
        part = T.synthetic_solenoidal_ic(n, block=blk)
        assert np.max(np.abs(part - full[(slice(None),) + blk])) <= 1e-13


def test_synthetic_field_is_the_same_function_on_every_grid():
    """cell centres of a 2x finer grid bracket the coarse ones: the field is a band-limited function of x, not of the grid"""
    a, b = T.synthetic_solenoidal_ic(16)[1], T.synthetic_solenoidal_ic(48)[1]
    # coarse centre (i + 1/2) h = fine centre (3 i + 1 + 1/2) h / 3
    assert np.max(np.abs(a - b[1::3, 1::3, 1::3])) <= 1e-12


@pytest.mark.parametrize("spectral", [False, True])
@pytest.mark.parametrize("spectrum", ["EXPONENTIAL", "KOLMOGOROV", "BOX"])
def test_hit_generator_is_pinned_to_the_reference(spectrum, spectral):
    from oracle.refharness import run_reference as rr
    if not rr.reference_available():
        pytest.skip("needs /root/reference (build container only)")
    if spectral and spectrum != "EXPONENTIAL":
        pytest.skip("the reference allows the spectral construction with the exponential spectrum only")
    rr._activate()
    from jaxfluids.turbulence.initialization import hit as ref_hit
    from jaxfluids.data_types.case_setup.initial_conditions import HITParameters
    import jax.numpy as jnp
    n, gamma, R = 16, 1.4, 4.4642857142857135
    par = HITParameters(1.0, 1.0, spectrum, 0.4, "IC1", 4, 8, spectral)
    np.random.seed(3)                                       # turb_init_manager.py:44-45
    mesh = [jnp.asarray(m * 1.0) for m in np.meshgrid(*(np.arange(n),) * 3, indexing="ij")]
    ref = np.asarray(ref_hit.initialize_hit(mesh, (1, 1, 1), gamma, R, par))
    mine = T.initialize_hit(n, gamma, R, energy_spectrum=spectrum, xi_0=4, xi_1=8, ma_target=0.4, T_ref=1.0, rho_ref=1.0,
                            is_velocity_spectral=spectral, random_seed=3)
    assert np.array_equal(ref, mine)


def test_shipped_hit_case_file_initial_condition_parses():
    """examples_3D/02_hit/HIT_decay.json's initial_condition / domain / EOS blocks select the generator (its numerical
    setup -- ALDM, viscous with a temperature-dependent viscosity -- is not this path: TGV numerics instead)."""
    from jaxfluids_b200.input_manager import InputManager
    import bench
    case = {
        "general": {"case_name": "HIT", "end_time": 5.0, "save_path": "./results", "save_dt": 0.1},
        "domain": {ax: {"cells": 16, "range": [0.0, 6.283185307179586]} for ax in "xyz"},
        "boundary_conditions": {f: {"type": "PERIODIC"} for f in port.FACES},
        "initial_condition": {"turbulent": {"case": "HIT", "random_seed": 0, "parameters": {
            "energy_spectrum": "EXPONENTIAL", "xi_0": 4, "ma_target": 0.4, "T_ref": 1.0, "rho_ref": 1.0, "ic_type": "IC1"}}},
        "material_properties": {"equation_of_state": {"model": "IdealGas", "specific_heat_ratio": 1.4,
                                                      "specific_gas_constant": 4.4642857142857135}},
    }
    case["domain"]["decomposition"] = {"split_x": 1, "split_y": 1, "split_z": 1}
    im = InputManager(case, bench.numerical_setup())
    tb = im.case_setup.initial_condition_setup["turbulent"]
    assert tb["energy_spectrum"] == "EXPONENTIAL" and tb["xi_1"] == 16 and tb["is_velocity_spectral"] is False
    bad = json.loads(json.dumps(case))
    bad["initial_condition"]["turbulent"]["parameters"]["ic_type"] = "IC3"
    prims = T.initialize_hit(16, 1.4, 4.4642857142857135, **{k: v for k, v in tb.items() if k != "case"})
    assert prims.shape == (5, 16, 16, 16) and abs(np.sqrt(np.mean((prims[1:4] ** 2).sum(0))) / 2.5 - 0.4) < 1e-12
    with pytest.raises(NotImplementedError):
        T.initialize_hit(16, 1.4, 1.0, **{**{k: v for k, v in tb.items() if k != "case"}, "ic_type": "IC3"})


def test_oracle_reproduces_the_hit_fixture_bit_for_bit():
    g, case, num, s = _fixture_setup()
    with np.errstate(all="ignore"):
        prims, cons = port.initialize(g["user"], s)
    dt = port.time_step_size(prims, s)
    assert dt == float(g["dt0"])
    assert np.array_equal(port.compute_rhs(prims, s), g["rhs_s0"])
    for i in range(len(g["dt"])):
        prims, cons, dt = port.step(prims, cons, dt, s)
        assert dt == g["dt"][i]
    assert np.array_equal(prims[(slice(None),) + s.interior], g["prims_n3"])


@pytest.mark.gpu
def test_gpu_matches_the_hit_fixture():
    from jaxfluids_b200.engine import BlockState
    from tests.test_gpu_parity import make_solver, dev, host
    g, case, num, s = _fixture_setup()
    with np.errstate(all="ignore"):
        prims, cons = port.initialize(g["user"], s)
    sol = make_solver(s)
    got = host(sol.compute_rhs(dev(prims)))
    strict = H.rel_linf(got, g["rhs_s0"])
    terms = H.rel_linf(got, g["rhs_s0"], scale=H.rhs_scales(prims, s))
    print(f"\nHIT-like 32^3: stage-0 rhs rel Linf strict {strict:.2e}, over the axis terms {terms:.2e}")
    assert terms <= H.TOL_RHS and strict <= 1e-10
    st = BlockState(sol, prims, cons)
    assert abs(st.dt.item() - float(g["dt0"])) <= 1e-14 * float(g["dt0"])
    sl = (slice(None),) + s.interior
    for i in range(len(g["dt"])):
        st.step()
        assert abs(st.dt.item() - g["dt"][i]) <= 1e-12 * g["dt"][i]
        tot = np.array([host(st.conservatives)[sl][v].sum() for v in range(5)])
        ref = g["totals"][i]
        scale = np.maximum(np.abs(ref), 1e-3 * np.max(np.abs(ref)))
        assert np.max(np.abs(tot - ref) / scale) <= H.TOL_TOTALS * 20
    assert H.rel_linf(host(st.primitives)[sl], g["prims_n3"]) <= 1e-12
