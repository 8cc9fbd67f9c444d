"""CPU: the device numerics (jaxfluids_b200/csrc/numerics.cuh) compiled for the host.

(1) Without FMA contraction the device functions are bit-identical to the reference fixtures, i.e.
    they perform the reference's operations in the reference's order.
(2) With FMA contraction (what nvcc does by default) they stay within the north-star tolerance in the
    conditioning-aware norm, and the deviation in the cancelled-total norm shows why that norm is the
    right one (it measures the reference formula's conditioning at low Mach, not the implementation).
"""
import numpy as np
import pytest

from oracle import port
from tests import helpers as H
from tests import hostsim


def total_rhs(prims, s, fma, reference_order=False):
    tot = 0.0
    for a in s.active:
        tot = tot + hostsim.rhs_axis(prims, a, s, fma=fma, reference_order=reference_order)
    return tot


@pytest.mark.parametrize("name", H.golden_names(dissipative=False))
def test_device_functions_without_fma_are_bit_identical_to_reference(name):
    g, case, num = H.load_golden(name)
    s = H.setup_from_json(case, num)
    for a in s.active:
        got = hostsim.rhs_axis(g["prims0_halo"], a, s, fma=False, reference_order=True)
        assert np.array_equal(0.0 + got, g[f"rhs_axis{a}"])


@pytest.mark.parametrize("reference_order", [True, False])
@pytest.mark.parametrize("name", H.golden_names(dissipative=False))
def test_device_functions_with_fma_within_tolerance(name, reference_order):
    """Both evaluations (reference order / production re-association), FMA-contracted, every stage of
    the first step of every reference fixture."""
    g, case, num = H.load_golden(name)
    s = H.setup_from_json(case, num)
    for k in range(port.RK[s.integrator]["stages"]):
        p_in = g["prims0_halo"] if k == 0 else g[f"prims_s{k-1}"]
        p_in = np.nan_to_num(p_in, nan=1.0, posinf=1.0, neginf=1.0)
        got = total_rhs(p_in, s, fma=True, reference_order=reference_order)
        assert H.rel_linf(got, g[f"rhs_s{k}"], scale=H.rhs_scales(p_in, s)) <= H.TOL_RHS


def test_cancelled_total_norm_measures_conditioning_not_implementation():
    """TGV (Mach 0.1): the reference's own arithmetic with vs without FMA differs by >1e-12 relative to
    the cancelled total, and by <1e-12 relative to the summed terms."""
    g, case, num = H.load_golden("tgv16_sym_char_hllc_rk3")
    s = H.setup_from_json(case, num)
    a = total_rhs(g["prims0_halo"], s, fma=True, reference_order=True)
    b = total_rhs(g["prims0_halo"], s, fma=False, reference_order=True)
    assert H.rel_linf(a, b) > 1e-12
    assert H.rel_linf(a, b, scale=H.rhs_scales(g["prims0_halo"], s)) < 1e-12


@pytest.mark.parametrize("name", H.golden_names(dissipative=False))
def test_marching_variant_with_carried_weights(name):
    """sweep_strided's face_flux_carry (cell-centred weights of the as-is fields carried between
    consecutive faces) against the reference fixtures, every axis."""
    g, case, num = H.load_golden(name)
    s = H.setup_from_json(case, num)
    p_in = g["prims0_halo"]
    scales = H.rhs_scales(p_in, s)
    for a in s.active:
        got = hostsim.rhs_axis_march(p_in, a, s, fma=True)
        assert H.rel_linf(got, g[f"rhs_axis{a}"], scale=scales) <= H.TOL_RHS


@pytest.mark.parametrize("stencil", ["WENO5-Z", "WENO5-JS"])
@pytest.mark.parametrize("sig", ["EINFELDT", "ARITHMETIC", "RUSANOV", "DAVIS", "TORO"])
def test_signal_speeds_and_stencils_host_simulated(sig, stencil):
    """godunov.signal_speed x reconstruction_stencil: the device functions compiled for the host -- production
    evaluation (with FMA) and its marching form within 1e-12 of the pinned oracle, the reference-order evaluation
    (no FMA) bit-identical to it."""
    for cells, recon, bc in [((48, 1, 1), "CHAR-PRIMITIVE", "ZEROGRADIENT"), ((14, 18, 1), "PRIMITIVE", "PERIODIC"),
                             ((8, 10, 12), "CHAR-PRIMITIVE", "SYMMETRY")]:
        s = H.make_setup(cells, bc=bc, recon=recon, stencil=stencil)
        s.signal_speed = sig
        prims, cons = port.initialize(H.smooth_ic(s, seed=4, amp=0.2), s)
        scales = H.rhs_scales(prims, s)
        for a in s.active:
            ref = port.rhs_axis(prims, a, s)
            assert H.rel_linf(hostsim.rhs_axis(prims, a, s, fma=True), ref, scale=scales) <= H.TOL_RHS
            assert H.rel_linf(hostsim.rhs_axis_march(prims, a, s, fma=True), ref, scale=scales) <= H.TOL_RHS
            assert np.array_equal(hostsim.rhs_axis(prims, a, s, fma=False, reference_order=True), ref)


@pytest.mark.parametrize("tag", ["lv0", "lv1"])
def test_interpolation_limiter_host_simulated(tag):
    """limit_interpolation (numerics.cuh) on the reference's near-vacuum fixture, where it fires thousands of times."""
    import json, os
    g = np.load(os.path.join(H.GOLDEN, "special", "limiter_riemann2d_20x24.npz"))
    case, num = json.loads(str(g[f"case_json_{tag}"])), json.loads(str(g[f"num_json_{tag}"]))
    s = H.setup_from_json(case, num)
    with np.errstate(all="ignore"):
        p = g[f"prims_halo_{tag}"]
        tot = 0.0
        for a in s.active:
            tot = tot + hostsim.rhs_axis(p, a, s, fma=True)
        assert H.rel_linf(tot, g[f"rhs_{tag}"], scale=H.rhs_scales(p, s)) <= H.TOL_RHS
        tot0 = 0.0
        for a in s.active:
            tot0 = tot0 + hostsim.rhs_axis(p, a, s, fma=False, reference_order=True)
        assert np.array_equal(tot0, g[f"rhs_{tag}"], equal_nan=True)


@pytest.mark.parametrize("sig", ["EINFELDT", "ARITHMETIC", "RUSANOV", "DAVIS", "TORO"])
def test_hll_host_simulated(sig):
    """riemann_solver = HLL (HLL.py) with every signal speed: production evaluation within 1e-12 of the pinned
    oracle, reference-order evaluation bit-identical."""
    for cells, recon, bc in [((48, 1, 1), "CHAR-PRIMITIVE", "ZEROGRADIENT"), ((14, 18, 1), "PRIMITIVE", "PERIODIC"),
                             ((8, 10, 12), "CHAR-PRIMITIVE", "SYMMETRY")]:
        s = H.make_setup(cells, bc=bc, recon=recon, riemann="HLL")
        s.signal_speed = sig
        prims, cons = port.initialize(H.smooth_ic(s, seed=8, amp=0.2), s)
        scales = H.rhs_scales(prims, s)
        for a in s.active:
            ref = port.rhs_axis(prims, a, s)
            assert H.rel_linf(hostsim.rhs_axis(prims, a, s, fma=True), ref, scale=scales) <= H.TOL_RHS
            assert H.rel_linf(hostsim.rhs_axis_march(prims, a, s, fma=True), ref, scale=scales) <= H.TOL_RHS
            assert np.array_equal(hostsim.rhs_axis(prims, a, s, fma=False, reference_order=True), ref)


@pytest.mark.parametrize("tag", ["simple", "nasa", "simple_cellsize", "nasa_interp"])
def test_flux_limiter_host_simulated(tag):
    """positivity/flux_limiter (limiter_flux.py:146-330) on the reference's fixture where hundreds of faces switch to
    the first-order flux: reference-order evaluation bit-identical to the reference's rhs; production evaluation
    within 1e-12 with the SAME faces switched (a flipped face would show up as an O(1) difference)."""
    import json
    import os
    g = np.load(os.path.join(H.GOLDEN, "special", "flux_limiter_riemann2d_20x24.npz"))
    s = H.setup_from_json(json.loads(str(g[f"case_json_{tag}"])), json.loads(str(g[f"num_json_{tag}"])))
    prims, dt = g[f"prims_halo_{tag}"], float(g[f"dt_{tag}"])
    assert s.flux_limiter in ("SIMPLE", "NASA")
    scales = H.rhs_scales(prims, s)
    with np.errstate(all="ignore"):
        ref_total = 0.0
        for a in s.active:
            ref = port.rhs_axis(prims, a, s, g[f"cons_halo_{tag}"], dt)
            ref_total = ref_total + ref
            assert np.array_equal(hostsim.rhs_axis(prims, a, s, fma=False, reference_order=True, dt=dt), ref)
            assert H.rel_linf(hostsim.rhs_axis(prims, a, s, fma=True, dt=dt), ref, scale=scales) <= H.TOL_RHS
            assert H.rel_linf(hostsim.rhs_axis_march(prims, a, s, fma=True, dt=dt), ref, scale=scales) <= H.TOL_RHS
    assert np.array_equal(ref_total, g[f"rhs_{tag}"])
