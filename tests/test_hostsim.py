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
