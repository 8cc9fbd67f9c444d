"""Registration of the XLA FFI handlers (csrc/jxf_xla_ffi.cc) with jax.ffi -- the binding a JAX-Fluids
maintainer would use to call this library from inside the reference's jitted step.

Optional and NOT exercised in this repository: `jax` / `jaxlib` are not installed in the build image, so the
handler library cannot be compiled or tested here.  Importing this module without jax raises ImportError with
that explanation; nothing else in the package imports it.
"""
from __future__ import annotations

import ctypes
import os

try:
    import jax
    import jax.numpy as jnp
except Exception as exc:  # pragma: no cover - jax is absent in this image
    raise ImportError("jaxfluids_b200.jax_ffi needs jax/jaxlib (and libjxf_b200_ffi.so built from "
                      "csrc/jxf_xla_ffi.cc against jaxlib's XLA FFI headers); the ctypes path "
                      "(jaxfluids_b200.engine) needs neither") from exc

TARGETS = {"jxf_compute_rhs": "JxfComputeRhs", "jxf_stage": "JxfStage", "jxf_halo_fill": "JxfHaloFill",
           "jxf_time_step": "JxfTimeStep", "jxf_integrate_stage": "JxfIntegrateStage"}


def register(ffi_library_path: str | None = None):
    """jax.ffi.register_ffi_target for every handler symbol; returns the loaded library."""
    here = os.path.dirname(os.path.abspath(__file__))
    lib = ctypes.CDLL(ffi_library_path or os.path.join(here, "lib", "libjxf_b200_ffi.so"))
    for target, symbol in TARGETS.items():
        jax.ffi.register_ffi_target(target, jax.ffi.pycapsule(getattr(lib, symbol)), platform="CUDA")
    return lib


def compute_rhs(handle: int, primitives, interior_shape):
    """SpaceSolver.compute_rhs (space_solver.py:151) as an XLA custom call."""
    out = jax.ShapeDtypeStruct((5,) + tuple(interior_shape), jnp.float64)
    return jax.ffi.ffi_call("jxf_compute_rhs", out)(primitives, handle=int(handle))


def stage(handle: int, k: int, primitives, conservatives, conservatives_n, dt, red, interior_shape, reduce: bool):
    """One fused RK stage (simulation_manager.py:796-963): -> (primitives, conservatives, rhs, red)."""
    like = lambda a: jax.ShapeDtypeStruct(a.shape, a.dtype)
    rhs = jax.ShapeDtypeStruct((5,) + tuple(interior_shape), jnp.float64)
    return jax.ffi.ffi_call("jxf_stage", (like(primitives), like(conservatives), rhs, like(red)),
                            input_output_aliases={4: 3})(
        primitives, conservatives, conservatives_n, dt, red, handle=int(handle), stage=int(k), reduce=int(bool(reduce)),
        fill_halo=1)
