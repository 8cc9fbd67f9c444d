"""Host simulation of the device numerics (see host_numerics.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libhost_numerics.so")
SRC = os.path.join(HERE, "host_numerics.cpp")
HDR = os.path.join(os.path.dirname(os.path.dirname(HERE)), "jaxfluids_b200", "csrc", "numerics.cuh")


def load(fma=True, reference_order=False):
    """fma: contract a*b+c like nvcc does; reference_order: compile the JXF_REFERENCE_ORDER evaluation
    (the reference's operation order) instead of the production evaluation."""
    so = SO.replace(".so", f"_{'fma' if fma else 'nofma'}_{'ref' if reference_order else 'prod'}.so")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        flags = ["-O2", "-mfma", "-ffp-contract=fast"] if fma else ["-O2", "-ffp-contract=off"]
        if reference_order:
            flags.append("-DJXF_REFERENCE_ORDER")
        subprocess.run(["g++", "-shared", "-fPIC", "-std=c++17", "-I/usr/local/cuda/include"] + flags +
                       ["-o", so, SRC], check=True)
    lib = C.CDLL(so)
    lib.face_flux_host.restype = C.c_int
    lib.face_flux_host.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_long, C.c_double, C.c_void_p, C.c_int,
                                   C.c_double, C.c_double, C.c_double]
    return lib


STENCIL_IDS = {"WENO5-Z": 0, "WENO5-JS": 1, "WENO1": 2, "WENO3-JS": 3, "WENO3-Z": 4, "TENO5": 5, "WENO6-CU": 6,
               "KOREN": 7, "MC": 8, "MINMOD": 9, "SUPERBEE": 10, "VANALBADA": 11, "VANLEER": 12, "WENO3-N": 13,
               "CENTRAL2": 14, "TENO6": 15, "TENO5-A": 16, "TENO6-A": 17}   # JXF_STENCIL_*


def _stencil_bits(st):
    """stencil id in the option word: bits 11-14 + the fifth id bit at bit 22 (jxf_b200.cu stencil_bits)."""
    return ((st & 15) << 11) | ((st >> 4) << 22)


VARIABLE_IDS = {"PRIMITIVE": 0, "CHAR-PRIMITIVE": 1, "CONSERVATIVE": 2, "CHAR-CONSERVATIVE": 3}   # JXF_RECON_*


def _generic(s):
    """Whether base_args / dispatch_recon (jxf_b200.cu) route this setup to the generic kernel instantiations."""
    return (s.convective_solver == "FLUX-SPLITTING" or STENCIL_IDS[s.stencil] >= 2 or VARIABLE_IDS[s.recon] >= 2 or
            s.frozen_state == "ROE")


def _recon_id(s):
    """The kernels' RECON template parameter: variable + 2 * min(stencil, STENCIL_GENERIC) (dispatch_recon)."""
    if s.convective_solver == "FLUX-SPLITTING":          # always the generic instantiation (dispatch_recon)
        return 4
    var = VARIABLE_IDS[s.recon]
    if var >= 2 or s.frozen_state == "ROE":               # the conservative forms / the ROE frozen state: generic too
        return 4 + (var & 1)
    return var + 2 * min(STENCIL_IDS[s.stencil], 2)


def _riemann_id(s):
    """The kernels' RIEMANN template parameter: HLLC (0), or the RUSANOV instantiation (1) for everything else
    (dispatch_riemann / riemann_template_of, jxf_b200.cu)."""
    return 1 if s.convective_solver == "FLUX-SPLITTING" else {"HLLC": 0}.get(s.riemann, 1)


def _opt(s):
    """face_flux `opt` (numerics.cuh): limiter mode | signal speed << 4 | HLL << 8 | flux limiter << 9 |
    generic stencil id << 11, as base_args (jxf_b200.cu) packs it."""
    lim = (2 if s.limit_velocity else 1) if s.is_interpolation_limiter else 0
    sig = {"EINFELDT": 0, "ARITHMETIC": 1, "RUSANOV": 2, "DAVIS": 3, "TORO": 4}[s.signal_speed]
    fl = {None: 0, "SIMPLE": 1, "NASA": 2}[s.flux_limiter]
    st = STENCIL_IDS[s.stencil]
    alt = {"HLLC-LM": 1, "AUSMP": 2}.get(s.riemann, 0)              # RIEMANN_ALT_* (ride on the RUSANOV instantiations)
    roe = 1 if s.frozen_state == "ROE" else 0
    if s.convective_solver == "FLUX-SPLITTING":          # stencil id (all of them) + eigenvalue choice + frozen state
        return _stencil_bits(st) | ({"ROE": 1, "CLLF": 2, "LLF": 3}[s.flux_splitting] << 17) | (roe << 21)
    gen = _stencil_bits(st) | (VARIABLE_IDS[s.recon] << 19) | (roe << 21) if _generic(s) else 0
    return lim | (sig << 4) | ((1 if s.riemann == "HLL" else 0) << 8) | (fl << 9) | gen | (alt << 15)


def _flux_limiter_args(s, axis, dt):
    """(dt, 1/dx, sigma) of numerics.cuh FluxLimArgs, as base_args (jxf_b200.cu) fills them."""
    sigma = float(len(s.active))
    if s.flux_partition == "CELLSIZE":
        sigma = sum(s.inv_dx[a] for a in s.active) / s.inv_dx[axis]
    if s.flux_limiter and dt is None:
        raise ValueError("flux limiter: dt required")
    return float(dt or 0.0), float(s.inv_dx[axis]), sigma


def rhs_axis_march(prims, axis, s, fma=True, dt=None):
    """rhs_axis through the MARCHING variant of the device functions (weights of the as-is fields carried
    from face to face along the sweep axis, as sweep_strided does)."""
    from oracle import port
    lib = load(fma, False)
    lib.face_flux_march_host.restype = C.c_int
    lib.face_flux_march_host.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_long, C.c_long, C.c_double, C.c_void_p,
                                         C.c_int, C.c_double, C.c_double, C.c_double]
    w = np.stack(port._window(prims, axis, s), axis=-1)          # (5, X, Y, Z faces..., 6)
    w = np.moveaxis(w, 0, -2)                                    # (fx, fy, fz, 5, 6)
    w = np.moveaxis(w, axis, 2)                                  # sweep axis last of the three -> sequences
    shp = w.shape[:3]
    w = np.ascontiguousarray(w.reshape(-1, 5, 6))
    out = np.empty((w.shape[0], 5))
    rc = lib.face_flux_march_host(axis, _recon_id(s), _riemann_id(s),
                                  w.ctypes.data, shp[0] * shp[1], shp[2], s.gamma, out.ctypes.data, _opt(s),
                                  *_flux_limiter_args(s, axis, dt))
    assert rc == 0
    f = out.reshape(shp + (5,))
    f = np.moveaxis(f, 2, axis)                                  # back to (fx, fy, fz, 5)
    f = np.moveaxis(f, -1, 0)
    lo = [slice(None)] * 4
    hi = [slice(None)] * 4
    lo[1 + axis] = slice(None, -1)
    hi[1 + axis] = slice(1, None)
    return s.inv_dx[axis] * (f[tuple(lo)] - f[tuple(hi)])


def rhs_axis(prims, axis, s, fma=True, reference_order=False, dt=None):
    """Same contract as oracle.port.rhs_axis, computed with the device functions on the host."""
    from oracle import port
    lib = load(fma, reference_order)
    w = np.stack(port._window(prims, axis, s), axis=-1)          # (5, faces..., 6)
    w = np.moveaxis(w, 0, -2)                                    # (faces..., 5, 6)
    shp = w.shape[:-2]
    w = np.ascontiguousarray(w.reshape(-1, 5, 6))
    out = np.empty((w.shape[0], 5))
    rc = lib.face_flux_host(axis, _recon_id(s), _riemann_id(s),
                            w.ctypes.data, w.shape[0], s.gamma, out.ctypes.data, _opt(s),
                            *_flux_limiter_args(s, axis, dt))
    assert rc == 0
    f = np.moveaxis(out.reshape(shp + (5,)), -1, 0)
    lo = [slice(None)] * 4
    hi = [slice(None)] * 4
    lo[1 + axis] = slice(None, -1)
    hi[1 + axis] = slice(1, None)
    return s.inv_dx[axis] * (f[tuple(lo)] - f[tuple(hi)])
