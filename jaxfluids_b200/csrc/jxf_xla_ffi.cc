// jxf_xla_ffi.cc -- XLA FFI handlers over the C ABI of include/jxf_b200.h, for
//   jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(lib.<Symbol>), platform="CUDA")
// on the reference side (JAX-Fluids' SpaceSolver / SimulationManager).
//
// NOT BUILT IN THIS REPOSITORY'S IMAGE: it needs the XLA FFI headers that ship with jaxlib
// (jaxlib/include/xla/ffi/api/{c_api.h,ffi.h}); jax / jaxlib are not installed here and there is no network.
// Build where jaxlib is available:
//   nvcc -std=c++17 -shared -Xcompiler -fPIC -I$(python -c "import jaxlib,os;print(os.path.join(os.path.dirname(jaxlib.__file__),'include'))") \
//        -Iinclude jaxfluids_b200/csrc/jxf_xla_ffi.cc -Ljaxfluids_b200/lib -ljxf_b200 -o libjxf_b200_ffi.so
// Every handler only forwards to the C ABI: no arithmetic lives here, so the parity evidence of the C ABI
// (tests/, driven through ctypes) carries over.  The solver handle (jxf_create) is created once on the Python side
// through ctypes and passed as an i64 attribute.
#ifdef JXF_HAVE_XLA_FFI

#include <cuda_runtime.h>
#include <cstdint>

#include "xla/ffi/api/c_api.h"
#include "xla/ffi/api/ffi.h"

#include "jxf_b200.h"

namespace ffi = xla::ffi;
using F64 = ffi::Buffer<ffi::F64>;
using F64Out = ffi::ResultBuffer<ffi::F64>;

static ffi::Error status(int rc) {
  return rc >= 0 ? ffi::Error::Success() : ffi::Error::Internal(jxf_last_error());
}

// SpaceSolver.compute_rhs (solvers/space_solver.py:151): prims -> rhs
static ffi::Error ComputeRhsImpl(cudaStream_t stream, int64_t handle, F64 prims, F64Out rhs) {
  return status(jxf_compute_rhs(reinterpret_cast<jxf_handle>(handle), prims.typed_data(), rhs->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(JxfComputeRhs, ComputeRhsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("handle")
                                  .Arg<F64>()
                                  .Ret<F64>());

// One fused RK stage (simulation_manager.py:770-1047).  red is updated in place: alias operand 4 with result 3
// (input_output_aliases={4: 3}) on the Python side.
static ffi::Error StageImpl(cudaStream_t stream, int64_t handle, int32_t stage, int32_t reduce, int32_t fill_halo,
                            F64 prims_in, F64 cons_in, F64 cons_n, F64 dt, F64 red_in, F64Out prims_out,
                            F64Out cons_out, F64Out rhs_scratch, F64Out red_out) {
  (void)red_in;
  return status(jxf_stage(reinterpret_cast<jxf_handle>(handle), stage, prims_in.typed_data(), prims_out->typed_data(),
                          cons_in.typed_data(), cons_n.typed_data(), cons_out->typed_data(), rhs_scratch->typed_data(),
                          dt.typed_data(), red_out->typed_data(), reduce, fill_halo, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(JxfStage, StageImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("handle")
                                  .Attr<int32_t>("stage")
                                  .Attr<int32_t>("reduce")
                                  .Attr<int32_t>("fill_halo")
                                  .Arg<F64>()   // prims_in
                                  .Arg<F64>()   // cons_in
                                  .Arg<F64>()   // cons_n
                                  .Arg<F64>()   // dt (device scalar)
                                  .Arg<F64>()   // red (aliased with the last result)
                                  .Ret<F64>()   // prims_out
                                  .Ret<F64>()   // cons_out
                                  .Ret<F64>()   // rhs scratch (interior only)
                                  .Ret<F64>()); // red

// HaloManager.perform_halo_update_material, outer boundaries (halos/halo_manager.py:146): in place -> alias both
static ffi::Error HaloFillImpl(cudaStream_t stream, int64_t handle, F64 prims_in, F64 cons_in, F64Out prims, F64Out cons) {
  if (prims->typed_data() != prims_in.typed_data() || cons->typed_data() != cons_in.typed_data())
    return ffi::Error::InvalidArgument("jxf_halo_fill works in place: pass input_output_aliases={0: 0, 1: 1}");
  return status(jxf_halo_fill(reinterpret_cast<jxf_handle>(handle), prims->typed_data(), cons->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(JxfHaloFill, HaloFillImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("handle")
                                  .Arg<F64>()
                                  .Arg<F64>()
                                  .Ret<F64>()
                                  .Ret<F64>());

// compute_time_step_size + positivity info (time_step_size.py:15, positivity_handler.py:245): prims -> (dt, info[3])
static ffi::Error TimeStepImpl(cudaStream_t stream, int64_t handle, F64 prims, F64Out dt, F64Out info, F64Out red) {
  jxf_handle h = reinterpret_cast<jxf_handle>(handle);
  int rc = jxf_reduce_reset(h, red->typed_data(), stream);
  if (rc == 0) rc = jxf_reduce(h, prims.typed_data(), red->typed_data(), stream);
  if (rc == 0) rc = jxf_finish_step(h, red->typed_data(), dt->typed_data(), nullptr, info->typed_data(), stream);
  return status(rc);
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(JxfTimeStep, TimeStepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("handle")
                                  .Arg<F64>()
                                  .Ret<F64>()    // dt (1)
                                  .Ret<F64>()    // info (3): max sum(|u_i| + c), min rho, min p
                                  .Ret<F64>());  // scratch (3)

// TimeIntegrator.perform_stage_integration (time_integrator.py:108), stand-alone
static ffi::Error IntegrateStageImpl(cudaStream_t stream, int64_t handle, int32_t stage, double dt, F64 cons, F64 cons_n,
                                     F64 rhs, F64Out cons_out) {
  return status(jxf_integrate_stage(reinterpret_cast<jxf_handle>(handle), stage, cons.typed_data(), cons_n.typed_data(),
                                    rhs.typed_data(), dt, cons_out->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(JxfIntegrateStage, IntegrateStageImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("handle")
                                  .Attr<int32_t>("stage")
                                  .Attr<double>("dt")
                                  .Arg<F64>()
                                  .Arg<F64>()
                                  .Arg<F64>()
                                  .Ret<F64>());

#endif  // JXF_HAVE_XLA_FFI
