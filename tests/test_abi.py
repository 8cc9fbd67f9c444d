"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/jxf_b200.h declares.
Only host-side entry points are called here (no kernel launches without a GPU)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from jaxfluids_b200 import _lib
from tests.helpers import ROOT


def declared_symbols():
    with open(os.path.join(ROOT, "include", "jxf_b200.h")) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(jxf_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported(built_library):
    lib = C.CDLL(built_library)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/jxf_b200.h but not exported"
    assert set(names) == set(_lib.EXPORTS), set(names) ^ set(_lib.EXPORTS)


def test_library_contains_sm100a_code(built_library):
    out = subprocess.run(["cuobjdump", "-lelf", built_library], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def _cfg(**kw):
    c = _lib.JxfConfig()
    n = kw.get("n", (16, 16, 16))
    for i in range(3):
        c.n[i] = n[i]
        c.inv_dx[i] = 1.0
    c.nh = kw.get("nh", 5)
    c.dx_min, c.gamma, c.cfl, c.fixed_dt = 1.0, kw.get("gamma", 1.4), 0.5, 0.0
    c.recon, c.riemann, c.signal_speed, c.integrator = kw.get("recon", 1), kw.get("riemann", 0), kw.get("sig", 0), kw.get("integ", 2)
    bc = kw.get("bc", [1] * 6)
    for i in range(6):
        c.bc[i] = bc[i]
    return c


def test_create_destroy_and_sizes(built_library):
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.jxf_create(C.byref(_cfg()), C.byref(h)) == 0
    assert lib.jxf_field_elems(h) == 5 * 26 ** 3
    assert lib.jxf_rhs_elems(h) == 5 * 16 ** 3
    assert lib.jxf_num_stages(h) == 3
    assert lib.jxf_face_slab_elems(h, 0) == 5 * 5 * 16 * 16
    assert lib.jxf_destroy(h) == 0
    h = C.c_void_p()
    assert lib.jxf_create(C.byref(_cfg(n=(200, 1, 1), bc=[3, 3, 0, 0, 0, 0], integ=1)), C.byref(h)) == 0
    assert lib.jxf_field_elems(h) == 5 * 210
    assert lib.jxf_num_stages(h) == 2
    lib.jxf_destroy(h)


@pytest.mark.parametrize("kw,code,frag", [
    (dict(nh=2), -1, "halo_cells"),
    (dict(recon=4), -2, "reconstruction_variable"),
    (dict(riemann=5), -2, "riemann_solver"),
    (dict(sig=9), -2, "signal_speed"),
    (dict(integ=4), -2, "integrator"),
    (dict(bc=[7] * 6), -2, "boundary type"),
    (dict(gamma=0.9), -1, "gamma"),
    (dict(n=(16, 16, 16), bc=[1, 1, 1, 1, 0, 0]), -1, "INACTIVE"),
])
def test_create_error_convention(built_library, kw, code, frag):
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.jxf_create(C.byref(_cfg(**kw)), C.byref(h))
    assert rc == code
    assert frag in lib.jxf_last_error().decode()
    with pytest.raises(_lib.JxfError):
        _lib.check(rc)


def test_null_arguments_are_rejected(built_library):
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.jxf_create(C.byref(_cfg()), C.byref(h)) == 0
    assert lib.jxf_compute_rhs(h, None, None, None) == -1
    assert lib.jxf_sweep(h, 5, None, None, 0, None) == -1
    assert lib.jxf_stage(h, 0, None, None, None, None, None, None, None, None, 0, 1, None) == -1
    assert lib.jxf_halo_fill(h, None, None, None) == -1
    lib.jxf_destroy(h)


def _block_config(s):
    from jaxfluids_b200.engine import BlockConfig
    return BlockConfig(cells=s.cells, inv_dx=tuple(float(x) for x in s.inv_dx), dx_min=float(s.dx_min), gamma=s.gamma,
                       bc=s.bc, nh=s.nh, recon=s.recon, stencil=s.stencil, riemann=s.riemann, signal_speed=s.signal_speed,
                       convective_solver=s.convective_solver, flux_splitting=s.flux_splitting, frozen_state=s.frozen_state,
                       integrator=s.integrator, is_interpolation_limiter=s.is_interpolation_limiter,
                       limit_velocity=s.limit_velocity, flux_limiter=s.flux_limiter, flux_partition=s.flux_partition)


def test_host_simulation_uses_the_kernels_template_parameters_and_option_word(built_library):
    """tests/hostsim validates the device functions for a (RECON, RIEMANN, option word) it derives from the setup; the
    kernels get theirs from dispatch_recon / dispatch_riemann / base_args (jxf_b200.cu).  jxf_debug_dispatch reports the
    latter -- no GPU needed -- and the two must agree for every option combination, or the host simulation would be
    validating something the GPU does not run."""
    import itertools
    from tests import helpers as H
    from tests import hostsim
    lib = _lib.load()
    stencils = list(_lib.STENCIL)
    n = 0
    combos = itertools.chain(
        itertools.product(["GODUNOV"], stencils, list(_lib.RECON), ["ARITHMETIC", "ROE"], list(_lib.RIEMANN),
                          ["EINFELDT", "TORO"], [(False, False, None), (True, True, "NASA")]),
        itertools.product(["FLUX-SPLITTING"], stencils, ["CHAR-PRIMITIVE"], ["ARITHMETIC", "ROE"], ["HLLC"], ["EINFELDT"],
                          [(False, False, None)]))
    for solver, stencil, recon, frozen, riemann, sig, (lim, limv, fluxlim) in combos:
        for fs in (["ROE", "CLLF", "LLF"] if solver == "FLUX-SPLITTING" else ["ROE"]):
            s = H.make_setup((12, 10, 8), bc="PERIODIC", recon=recon, riemann=riemann, stencil=stencil)
            s.convective_solver, s.flux_splitting, s.frozen_state, s.signal_speed = solver, fs, frozen, sig
            s.is_interpolation_limiter, s.limit_velocity, s.flux_limiter = lim, limv, fluxlim
            h = C.c_void_p()
            assert lib.jxf_create(C.byref(_block_config(s).to_c()), C.byref(h)) == 0, lib.jxf_last_error()
            r, m, o = C.c_int32(), C.c_int32(), C.c_int32()
            for axis in range(3):
                assert lib.jxf_debug_dispatch(h, axis, C.byref(r), C.byref(m), C.byref(o)) == 0
                assert (r.value, m.value, o.value) == (hostsim._recon_id(s), hostsim._riemann_id(s), hostsim._opt(s)), \
                    (solver, stencil, recon, frozen, riemann, sig, lim, fluxlim, fs)
            lib.jxf_destroy(h)
            n += 1
    assert n > 2000
