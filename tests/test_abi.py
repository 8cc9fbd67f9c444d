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
