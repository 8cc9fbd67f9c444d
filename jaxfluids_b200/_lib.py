"""ctypes binding of the C ABI declared in include/jxf_b200.h.

The shared library is built in-tree by `__graft_entry__.build()` (or
`python -m jaxfluids_b200.build`).  There is NO fallback: if the library is
missing, import of the compute path fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libjxf_b200.so")

# names declared in include/jxf_b200.h (kept in sync by tests/test_abi.py)
EXPORTS = (
    "jxf_last_error", "jxf_version", "jxf_create", "jxf_destroy", "jxf_field_elems", "jxf_rhs_elems",
    "jxf_num_stages", "jxf_compute_rhs", "jxf_sweep", "jxf_stage", "jxf_halo_fill", "jxf_prims_from_cons",
    "jxf_cons_from_prims", "jxf_reduce", "jxf_reduce_reset", "jxf_finish_step", "jxf_face_slab_elems",
    "jxf_pack_face", "jxf_unpack_face", "jxf_fp64_probe", "jxf_step_fused", "jxf_profile_enable", "jxf_profile_read", "jxf_debug_face_flux", "jxf_debug_dispatch", "jxf_debug_math", "jxf_stage_tail", "jxf_sweep_range", "jxf_integrate_stage", "jxf_halo_fill_edges", "jxf_dissipative_sweep", "jxf_temperature", "jxf_face_slab_elems_ext", "jxf_pack_face_ext", "jxf_unpack_face_ext", "jxf_bind_timestep",
    "jxf_face_slab_elems_n", "jxf_pack_face_n", "jxf_unpack_face_n", "jxf_rhs_slab_elems", "jxf_stage_inplace", "jxf_set_face_data",
    "jxf_set_peer_halo", "jxf_peer_signal", "jxf_peer_wait", "jxf_peer_export", "jxf_peer_import", "jxf_peer_release",
)

RECON = {"PRIMITIVE": 0, "CHAR-PRIMITIVE": 1, "CONSERVATIVE": 2, "CHAR-CONSERVATIVE": 3}
FROZEN_STATE = {"ARITHMETIC": 0, "ROE": 1}
# ids = include/jxf_b200.h JXF_STENCIL_*; >= 2: the generic (reference-order) kernel instantiations
STENCIL = {"WENO5-Z": 0, "WENO5-JS": 1, "WENO1": 2, "WENO3-JS": 3, "WENO3-Z": 4, "TENO5": 5, "WENO6-CU": 6,
           "KOREN": 7, "MC": 8, "MINMOD": 9, "SUPERBEE": 10, "VANALBADA": 11, "VANLEER": 12, "WENO3-N": 13,
           "CENTRAL2": 14, "TENO6": 15, "TENO5-A": 16, "TENO6-A": 17}
FLUX_LIMITER = {None: 0, False: 0, "SIMPLE": 1, "NASA": 2}
FLUX_PARTITION = {"UNIFORM": 0, "CELLSIZE": 1}
RIEMANN = {"HLLC": 0, "RUSANOV": 1, "HLL": 2, "HLLC-LM": 3, "AUSMP": 4}
SIGNAL = {"EINFELDT": 0, "ARITHMETIC": 1, "RUSANOV": 2, "DAVIS": 3, "TORO": 4}
CONVECTIVE_SOLVER = {"GODUNOV": 0, "FLUX-SPLITTING": 1}
FLUX_SPLITTING = {"ROE": 1, "CLLF": 2, "LLF": 3}
INTEGRATOR = {"EULER": 0, "RK2": 1, "RK3": 2, "RK2_LS4": 3}
BC = {"INACTIVE": 0, "PERIODIC": 1, "SYMMETRY": 2, "ZEROGRADIENT": 3, "NEIGHBOR": 4, "WALL": 5, "DIRICHLET": 6}
FACES = ("east", "west", "north", "south", "top", "bottom")


class JxfConfig(C.Structure):
    _fields_ = [
        ("n", C.c_int32 * 3),
        ("nh", C.c_int32),
        ("inv_dx", C.c_double * 3),
        ("dx_min", C.c_double),
        ("gamma", C.c_double),
        ("cfl", C.c_double),
        ("fixed_dt", C.c_double),
        ("recon", C.c_int32),
        ("riemann", C.c_int32),
        ("signal_speed", C.c_int32),
        ("integrator", C.c_int32),
        ("bc", C.c_int32 * 6),
        ("viscous_flux", C.c_int32),
        ("heat_flux", C.c_int32),
        ("viscous_heat_production", C.c_int32),
        ("stencil", C.c_int32),
        ("dynamic_viscosity", C.c_double),
        ("bulk_viscosity", C.c_double),
        ("thermal_conductivity", C.c_double),
        ("gas_constant", C.c_double),
        ("interpolation_limiter", C.c_int32),
        ("limit_velocity", C.c_int32),
        ("wall_velocity", (C.c_double * 3) * 6),
        ("dirichlet", (C.c_double * 5) * 6),
        ("volume_force", C.c_int32),
        ("no_convective_flux", C.c_int32),
        ("gravity", C.c_double * 3),
        ("flux_limiter", C.c_int32),
        ("flux_partition", C.c_int32),
        ("convective_solver", C.c_int32),
        ("flux_splitting", C.c_int32),
        ("frozen_state", C.c_int32),
    ]


class JxfError(RuntimeError):
    pass


_lib = None


def load():
    """Load libjxf_b200.so and declare the prototypes. Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    path = LIB_PATH
    variant = os.environ.get("JXF_LIB_VARIANT")
    if variant:
        path = LIB_PATH.replace(".so", f"_{variant}.so")
    if not os.path.exists(path):
        raise JxfError(
            f"{path} not found: the CUDA extension is not built. Run `python -c 'import __graft_entry__ as g; "
            "g.build()'` at the repo root. There is no CPU fallback for this path.")
    lib = C.CDLL(path)
    vp, dp, i32, i64 = C.c_void_p, C.c_void_p, C.c_int, C.c_int64
    lib.jxf_last_error.restype = C.c_char_p
    lib.jxf_last_error.argtypes = []
    lib.jxf_version.restype = i32
    lib.jxf_create.restype = i32
    lib.jxf_create.argtypes = [C.POINTER(JxfConfig), C.POINTER(vp)]
    lib.jxf_destroy.restype = i32
    lib.jxf_destroy.argtypes = [vp]
    for name in ("jxf_field_elems", "jxf_rhs_elems"):
        getattr(lib, name).restype = i64
        getattr(lib, name).argtypes = [vp]
    lib.jxf_num_stages.restype = i32
    lib.jxf_num_stages.argtypes = [vp]
    lib.jxf_compute_rhs.restype = i32
    lib.jxf_compute_rhs.argtypes = [vp, dp, dp, vp]
    lib.jxf_bind_timestep.restype = i32
    lib.jxf_bind_timestep.argtypes = [vp, dp]
    lib.jxf_sweep.restype = i32
    lib.jxf_sweep.argtypes = [vp, i32, dp, dp, i32, vp]
    lib.jxf_stage.restype = i32
    lib.jxf_stage.argtypes = [vp, i32, dp, dp, dp, dp, dp, dp, dp, dp, i32, i32, vp]
    lib.jxf_halo_fill.restype = i32
    lib.jxf_halo_fill.argtypes = [vp, dp, dp, vp]
    lib.jxf_prims_from_cons.restype = i32
    lib.jxf_prims_from_cons.argtypes = [vp, dp, dp, vp]
    lib.jxf_cons_from_prims.restype = i32
    lib.jxf_cons_from_prims.argtypes = [vp, dp, dp, vp]
    lib.jxf_reduce.restype = i32
    lib.jxf_reduce.argtypes = [vp, dp, dp, vp]
    lib.jxf_reduce_reset.restype = i32
    lib.jxf_reduce_reset.argtypes = [vp, dp, vp]
    lib.jxf_finish_step.restype = i32
    lib.jxf_finish_step.argtypes = [vp, dp, dp, dp, dp, vp]
    lib.jxf_face_slab_elems.restype = i64
    lib.jxf_face_slab_elems.argtypes = [vp, i32]
    lib.jxf_pack_face.restype = i32
    lib.jxf_pack_face.argtypes = [vp, i32, dp, dp, vp]
    lib.jxf_unpack_face.restype = i32
    lib.jxf_unpack_face.argtypes = [vp, i32, dp, dp, dp, vp]
    lib.jxf_fp64_probe.restype = i32
    lib.jxf_fp64_probe.argtypes = [dp, i32, C.POINTER(i64), vp]
    lib.jxf_step_fused.restype = i32
    lib.jxf_step_fused.argtypes = [vp, dp, dp, dp, dp, dp, dp, dp, dp, dp, i32, vp]
    lib.jxf_profile_enable.restype = i32
    lib.jxf_profile_enable.argtypes = [vp, i32]
    lib.jxf_profile_read.restype = i32
    lib.jxf_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(i64), C.POINTER(i64), i32]
    lib.jxf_stage_tail.restype = i32
    lib.jxf_stage_tail.argtypes = [vp, i32, i32, dp, dp, dp, dp, dp, dp, dp, dp, i32, i32, vp]
    lib.jxf_sweep_range.restype = i32
    lib.jxf_sweep_range.argtypes = [vp, i32, i32, i32, dp, dp, i32, vp]
    lib.jxf_integrate_stage.restype = i32
    lib.jxf_integrate_stage.argtypes = [vp, i32, dp, dp, dp, C.c_double, dp, vp]
    lib.jxf_halo_fill_edges.restype = i32
    lib.jxf_halo_fill_edges.argtypes = [vp, dp, dp, vp]
    lib.jxf_dissipative_sweep.restype = i32
    lib.jxf_dissipative_sweep.argtypes = [vp, i32, dp, dp, i32, vp]
    lib.jxf_temperature.restype = i32
    lib.jxf_temperature.argtypes = [vp, dp, dp, vp]
    lib.jxf_face_slab_elems_ext.restype = i64
    lib.jxf_face_slab_elems_ext.argtypes = [vp, i32, i32]
    lib.jxf_pack_face_ext.restype = i32
    lib.jxf_pack_face_ext.argtypes = [vp, i32, i32, dp, dp, vp]
    lib.jxf_unpack_face_ext.restype = i32
    lib.jxf_unpack_face_ext.argtypes = [vp, i32, i32, dp, dp, dp, vp]
    lib.jxf_face_slab_elems_n.restype = i64
    lib.jxf_face_slab_elems_n.argtypes = [vp, i32, i32, i32]
    lib.jxf_pack_face_n.restype = i32
    lib.jxf_pack_face_n.argtypes = [vp, i32, i32, i32, dp, dp, vp]
    lib.jxf_unpack_face_n.restype = i32
    lib.jxf_unpack_face_n.argtypes = [vp, i32, i32, i32, dp, dp, dp, vp]
    lib.jxf_rhs_slab_elems.restype = i64
    lib.jxf_rhs_slab_elems.argtypes = [vp, i32]
    lib.jxf_stage_inplace.restype = i32
    lib.jxf_stage_inplace.argtypes = [vp, i32, dp, dp, dp, dp, dp, i32, dp, dp, i32, i32, vp]
    lib.jxf_set_face_data.restype = i32
    lib.jxf_set_face_data.argtypes = [vp, i32, i32, dp, vp]
    lib.jxf_set_peer_halo.restype = i32
    lib.jxf_set_peer_halo.argtypes = [vp, i32, dp, dp]
    lib.jxf_peer_signal.restype = i32
    lib.jxf_peer_signal.argtypes = [vp, C.POINTER(C.c_void_p), i64, vp]
    lib.jxf_peer_export.restype = i32
    lib.jxf_peer_export.argtypes = [vp, vp, C.POINTER(C.c_int64)]
    lib.jxf_peer_import.restype = i32
    lib.jxf_peer_import.argtypes = [vp, i64, C.POINTER(C.c_void_p)]
    lib.jxf_peer_release.restype = i32
    lib.jxf_peer_release.argtypes = []
    lib.jxf_peer_wait.restype = i32
    lib.jxf_peer_wait.argtypes = [vp, dp, i32, i64, vp]
    lib.jxf_debug_math.restype = i32
    lib.jxf_debug_math.argtypes = [dp, i64, dp, vp]
    lib.jxf_debug_face_flux.restype = i32
    lib.jxf_debug_face_flux.argtypes = [i32, i32, i32, dp, i64, C.c_double, dp, vp]
    lib.jxf_debug_dispatch.restype = i32
    lib.jxf_debug_dispatch.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().jxf_last_error().decode("utf-8", "replace")
        raise JxfError(f"jxf error {rc}: {msg}")
