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


@pytest.mark.parametrize("name", H.golden_names(dissipative=False) + H.generic_golden_names())
def test_device_functions_without_fma_are_bit_identical_to_reference(name):
    g, case, num = H.load_golden(name)
    s = H.setup_from_json(case, num)
    for a in s.active:
        got = hostsim.rhs_axis(g["prims0_halo"], a, s, fma=False, reference_order=True)
        if s.recon in ("CONSERVATIVE", "CHAR-CONSERVATIVE") and s.convective_solver == "GODUNOV":
            # the Riemann solvers re-form the face conservatives from the face primitives (prims_from_cons of the
            # reconstructed conservatives), the reference hands them the reconstructed ones: rounding-level difference
            # (2e-14 at Mach 0.1, where the pressure is a small difference of the reconstructed conservatives)
            assert H.rel_linf(0.0 + got, g[f"rhs_axis{a}"], scale=H.rhs_scales(g["prims0_halo"], s)) <= 1e-13
        else:
            assert np.array_equal(0.0 + got, g[f"rhs_axis{a}"])


@pytest.mark.parametrize("reference_order", [True, False])
@pytest.mark.parametrize("name", H.golden_names(dissipative=False) + H.generic_golden_names())
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


@pytest.mark.parametrize("name", H.golden_names(dissipative=False) + H.generic_golden_names())
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


@pytest.mark.parametrize("variable", ["prim", "char"])
@pytest.mark.parametrize("stencil", H.GENERIC_STENCILS)
def test_generic_stencils_host_simulated(stencil, variable):
    """stencil_generic / reconstruct_generic (numerics.cuh) on the reference's shocked fixture, where the TENO cut-off
    and the slope limiters switch: without FMA bit-identical to the reference's rhs, with FMA (what nvcc emits) within
    1e-12, through face_flux and through the marching form."""
    import json
    import os
    g = np.load(os.path.join(H.GOLDEN, "special", "stencils_riemann2d_20x24.npz"))
    key = f"{stencil}_{variable}"
    s = H.setup_from_json(json.loads(str(g[f"case_json_{key}"])), json.loads(str(g[f"num_json_{key}"])))
    prims = g["prims_halo"]
    scales = H.rhs_scales(prims, s)
    tot_exact, tot_fma, tot_march = 0.0, 0.0, 0.0
    for a in s.active:
        tot_exact = tot_exact + hostsim.rhs_axis(prims, a, s, fma=False, reference_order=True)
        tot_fma = tot_fma + hostsim.rhs_axis(prims, a, s, fma=True)
        tot_march = tot_march + hostsim.rhs_axis_march(prims, a, s, fma=True)
    assert np.array_equal(tot_exact, g[f"rhs_{key}"])
    assert H.rel_linf(tot_fma, g[f"rhs_{key}"], scale=scales) <= H.TOL_RHS
    assert H.rel_linf(tot_march, g[f"rhs_{key}"], scale=scales) <= H.TOL_RHS


@pytest.mark.parametrize("stencil", ["WENO3-Z", "TENO5", "WENO6-CU", "VANALBADA", "TENO5-A", "TENO6-A"])
def test_generic_stencils_3d_all_axes_host_simulated(stencil):
    """Every sweep axis (the axis only enters through the velocity roles) and both reconstruction variables."""
    for cells, recon, bc in [((8, 10, 12), "CHAR-PRIMITIVE", "SYMMETRY"), ((9, 8, 10), "PRIMITIVE", "PERIODIC")]:
        s = H.make_setup(cells, bc=bc, recon=recon, stencil=stencil)
        prims, cons = port.initialize(H.smooth_ic(s, seed=11, amp=0.2), s)
        scales = H.rhs_scales(prims, s)
        for a in s.active:
            ref = port.rhs_axis(prims, a, s)
            assert np.array_equal(hostsim.rhs_axis(prims, a, s, fma=False, reference_order=True), ref)
            assert H.rel_linf(hostsim.rhs_axis(prims, a, s, fma=True), ref, scale=scales) <= H.TOL_RHS
            assert H.rel_linf(hostsim.rhs_axis_march(prims, a, s, fma=True), ref, scale=scales) <= H.TOL_RHS


def _fast_ic(s, seed, factor):
    """smooth_ic with the velocities scaled up: sub- and supersonic faces of both signs."""
    ic = H.smooth_ic(s, seed=seed, amp=0.2)
    ic[1:4] *= factor
    return ic


@pytest.mark.parametrize("riemann,sig", [("HLLC-LM", "EINFELDT"), ("HLLC-LM", "DAVIS"), ("HLLC-LM", "TORO"),
                                         ("AUSMP", "EINFELDT")])
def test_hllclm_and_ausmp_host_simulated(riemann, sig):
    """riemann_solver = HLLC-LM (HLLCLM.py) / AUSMP (AUSMP.py): riemann_other (numerics.cuh) without FMA bit-identical
    to the pinned oracle, with FMA within 1e-12 (face_flux and the marching form) -- on low-Mach states (the HLLC-LM
    limiter phi < 1) and on states with supersonic faces of both signs (the |M| >= 1 branches of AUSM+, sign(S_K))."""
    for cells, recon, bc, factor in [((48, 1, 1), "CHAR-PRIMITIVE", "ZEROGRADIENT", 4.0),
                                     ((14, 18, 1), "PRIMITIVE", "PERIODIC", 4.0),
                                     ((8, 10, 12), "CHAR-PRIMITIVE", "SYMMETRY", 0.05)]:
        s = H.make_setup(cells, bc=bc, recon=recon, riemann=riemann)
        s.signal_speed = sig
        prims, cons = port.initialize(_fast_ic(s, 4, factor), s)
        pi = prims[(slice(None),) + s.interior]
        mach = pi[1] / port.speed_of_sound(pi[4], pi[0], s.gamma)
        assert (mach.min() < -1.05 and mach.max() > 1.05) if factor > 1 else np.abs(mach).max() < 0.1
        scales = H.rhs_scales(prims, s)
        for a in s.active:
            ref = port.rhs_axis(prims, a, s)
            assert np.array_equal(hostsim.rhs_axis(prims, a, s, fma=False, reference_order=True), ref)
            assert H.rel_linf(hostsim.rhs_axis(prims, a, s, fma=True), ref, scale=scales) <= H.TOL_RHS
            assert H.rel_linf(hostsim.rhs_axis_march(prims, a, s, fma=True), ref, scale=scales) <= H.TOL_RHS


@pytest.mark.parametrize("fs,stencil", [("ROE", "WENO5-Z"), ("CLLF", "WENO6-CU"), ("LLF", "WENO5-JS"), ("ROE", "TENO5"),
                                        ("CLLF", "WENO3-Z"), ("LLF", "VANLEER")])
def test_flux_splitting_host_simulated(fs, stencil):
    """convective_solver = FLUX-SPLITTING (flux_splitting_scheme.py): flux_splitting_flux (numerics.cuh) without FMA
    bit-identical to the pinned oracle, with FMA within 1e-12 (face_flux and the marching form), every axis, sub- and
    supersonic states; the stencil ids of the two tuned WENO5 forms included (reference-order forms in the generic
    function)."""
    for cells, bc, factor in [((48, 1, 1), "ZEROGRADIENT", 4.0), ((14, 18, 1), "PERIODIC", 1.0), ((8, 10, 12), "SYMMETRY", 1.0)]:
        s = H.make_setup(cells, bc=bc, stencil=stencil)
        s.convective_solver, s.flux_splitting = "FLUX-SPLITTING", fs
        prims, cons = port.initialize(_fast_ic(s, 4, factor), s)
        scales = H.rhs_scales(prims, s)
        for a in s.active:
            ref = port.rhs_axis(prims, a, s)
            assert np.array_equal(hostsim.rhs_axis(prims, a, s, fma=False, reference_order=True), ref)
            assert H.rel_linf(hostsim.rhs_axis(prims, a, s, fma=True), ref, scale=scales) <= H.TOL_RHS
            assert H.rel_linf(hostsim.rhs_axis_march(prims, a, s, fma=True), ref, scale=scales) <= H.TOL_RHS


@pytest.mark.parametrize("recon,frozen,stencil,riemann", [
    ("CONSERVATIVE", "ARITHMETIC", "WENO5-Z", "HLLC"), ("CHAR-CONSERVATIVE", "ARITHMETIC", "WENO5-Z", "HLLC"),
    ("CHAR-CONSERVATIVE", "ROE", "TENO5", "HLL"), ("CHAR-PRIMITIVE", "ROE", "WENO5-Z", "HLLC"),
    ("CHAR-PRIMITIVE", "ROE", "WENO3-Z", "RUSANOV"), ("CONSERVATIVE", "ARITHMETIC", "VANLEER", "AUSMP"),
    ("FLUX-SPLITTING", "ROE", "WENO5-JS", "HLLC")])
def test_reconstruction_variables_and_roe_frozen_state_host_simulated(recon, frozen, stencil, riemann):
    """reconstruction_variable CONSERVATIVE / CHAR-CONSERVATIVE (reconstruct_conservative) and frozen_state ROE
    (frozen_state, numerics.cuh) through the generic path, sub- and supersonic states, every axis.  ROE with
    CHAR-PRIMITIVE or FLUX-SPLITTING is bit-identical to the pinned oracle without FMA; the conservative forms agree to
    rounding (the Riemann solvers re-form the face conservatives from the face primitives)."""
    for cells, bc, factor in [((48, 1, 1), "ZEROGRADIENT", 4.0), ((14, 18, 1), "PERIODIC", 1.0), ((8, 10, 12), "SYMMETRY", 1.0)]:
        fs = recon == "FLUX-SPLITTING"
        s = H.make_setup(cells, bc=bc, stencil=stencil, recon="CHAR-PRIMITIVE" if fs else recon, riemann=riemann)
        s.frozen_state = frozen
        if fs:
            s.convective_solver, s.flux_splitting = "FLUX-SPLITTING", "CLLF"
        prims, cons = port.initialize(_fast_ic(s, 4, factor), s)
        scales = H.rhs_scales(prims, s)
        for a in s.active:
            ref = port.rhs_axis(prims, a, s)
            exact = hostsim.rhs_axis(prims, a, s, fma=False, reference_order=True)
            if "CONSERVATIVE" in recon:
                assert H.rel_linf(exact, ref, scale=scales) <= 1e-14
            else:
                assert np.array_equal(exact, ref)
            assert H.rel_linf(hostsim.rhs_axis(prims, a, s, fma=True), ref, scale=scales) <= H.TOL_RHS
            assert H.rel_linf(hostsim.rhs_axis_march(prims, a, s, fma=True), ref, scale=scales) <= H.TOL_RHS
