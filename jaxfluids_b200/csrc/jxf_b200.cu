// jxf_b200.cu -- sm_100a kernels + C ABI (include/jxf_b200.h) of the convective path.
//
// Kernel family (one per sweep direction kind):
//   sweep_strided<A,..>: sweep axis A is NOT the contiguous axis. Lanes run along the
//       contiguous axis (coalesced), each thread marches along A with a rolling 6-cell
//       register window and keeps the previous face flux, so every face flux is computed once.
//   sweep_contig<A,..>:  sweep axis A IS the contiguous axis. Lanes = consecutive faces
//       of the flattened (row, face) sequence; the left face flux comes from lane-1 by
//       warp shuffle (lane 0: carry from the previous iteration).
// EPI=0 writes/accumulates the axis contribution into the interior-only rhs buffer
// (space_solver.py:597-599, :314); EPI=1 (last active axis of a stage) fuses the RK
// stage combination (RK3.py:49-60, time_integrator.py:57), primitive recovery
// (equation_manager.py:164-171) and the CFL / min-rho / min-p reductions
// (time_step_size.py:103-109, positivity_handler.py:246-247).
#include "plan.cuh"
#include <mutex>
#include <vector>

namespace jxf {

// ---------------------------------------------------------------------------
// halo fill: PERIODIC / SYMMETRY / ZEROGRADIENT face halos, cons recomputed
// (halos/outer/material.py:868-894, boundary_condition.py:563-595, :698-731)
// ---------------------------------------------------------------------------
struct HaloArgs {
  double* prims;
  double* cons;
  double gamma;
  int bc[6];
  double wall[6][3];
  double dirichlet[6][5];
  FaceData fd;             // per-face boundary data on top of the base rule (sweep_kernels.cuh), when has_fd
  int has_fd;
};

__global__ void __launch_bounds__(128) halo_fill_kernel(const Geom g, const HaloArgs a) {
  const int face = blockIdx.y;
  const int kind = a.bc[face];
  if (kind != JXF_BC_PERIODIC && kind != JXF_BC_SYMMETRY && kind != JXF_BC_ZEROGRADIENT && kind != JXF_BC_WALL &&
      kind != JXF_BC_DIRICHLET) return;
  const int ax = face >> 1;
  const bool hi = (face & 1) == 0;   // east, north, top
  // transverse axes: t2 is the faster (larger index) one
  const int t1 = (ax == 0) ? 1 : 0;
  const int t2 = (ax == 2) ? 1 : 2;
  const int n1 = g.n[t1], n2 = g.n[t2];
  const long long total = (long long)g.nh * n1 * n2;
  const int nh = g.nh, ext = g.ext[ax];
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total;
       q += (long long)gridDim.x * blockDim.x) {
    // halo layer l in increasing buffer index.  Faces of the CONTIGUOUS axis: the nh layers of a row are adjacent in
    // memory, so l runs fastest there (one 8 nh-byte segment per row instead of nh accesses a row pitch apart)
    int i1, i2, l;
    if (g.st[ax] == 1) {
      l = (int)(q % nh);
      const long long q1 = q / nh;
      i2 = (int)(q1 % n2);
      i1 = (int)(q1 / n2);
    } else {
      i2 = (int)(q % n2);
      const long long q1 = q / n2;
      i1 = (int)(q1 % n1);
      l = (int)(q1 / n1);
    }
    const int dst = hi ? (ext - nh + l) : l;
    int src;
    if (kind == JXF_BC_PERIODIC) src = hi ? (nh + l) : (ext - 2 * nh + l);
    else if (kind == JXF_BC_SYMMETRY || kind == JXF_BC_WALL) src = hi ? (ext - nh - 1 - l) : (2 * nh - 1 - l);
    else src = hi ? (ext - nh - 1) : nh;
    const long long tr = (long long)(i1 + g.off[t1]) * g.st[t1] + (long long)(i2 + g.off[t2]) * g.st[t2];
    const long long is = tr + (long long)src * g.st[ax];
    const long long id = tr + (long long)dst * g.st[ax];
    double p[5], c[5];
#pragma unroll
    for (int v = 0; v < 5; ++v) p[v] = a.prims[is + v * g.vst];
    if (kind == JXF_BC_SYMMETRY) p[1 + ax] = p[1 + ax] * -1.0;
    if (kind == JXF_BC_WALL) {
#pragma unroll
      for (int v = 1; v < 4; ++v) p[v] = 2 * a.wall[face][v - 1] - p[v];
    }
    if (kind == JXF_BC_DIRICHLET) {
#pragma unroll
      for (int v = 0; v < 5; ++v) p[v] = a.dirichlet[face][v];
    }
    if (a.has_fd) apply_face_data(a.fd, face, (long long)i1 * n2 + i2, (long long)n1 * n2, p);
    cons_from_prims(p, a.gamma, c);
#pragma unroll
    for (int v = 0; v < 5; ++v) {
      a.prims[id + v * g.vst] = p[v];
      a.cons[id + v * g.vst] = c[v];
    }
  }
}

// ---------------------------------------------------------------------------
// whole-buffer transforms, reductions, finish-step, pack/unpack
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) prims_from_cons_kernel(const double* __restrict__ cons, double* __restrict__ prims,
                                                              long long vst, double gamma) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < vst; i += (long long)gridDim.x * blockDim.x) {
    double c[5], p[5];
#pragma unroll
    for (int v = 0; v < 5; ++v) c[v] = cons[i + v * vst];
    prims_from_cons(c, gamma, p);
#pragma unroll
    for (int v = 0; v < 5; ++v) prims[i + v * vst] = p[v];
  }
}

__global__ void __launch_bounds__(256) cons_from_prims_kernel(const double* __restrict__ prims, double* __restrict__ cons,
                                                              long long vst, double gamma) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < vst; i += (long long)gridDim.x * blockDim.x) {
    double c[5], p[5];
#pragma unroll
    for (int v = 0; v < 5; ++v) p[v] = prims[i + v * vst];
    cons_from_prims(p, gamma, c);
#pragma unroll
    for (int v = 0; v < 5; ++v) cons[i + v * vst] = c[v];
  }
}

__global__ void __launch_bounds__(256) reduce_kernel(const Geom g, const double* __restrict__ prims, double* red,
                                                     double gamma, int active_mask) {
  const long long total = (long long)g.n[0] * g.n[1] * g.n[2];
  Red r;
  r.init();
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(q % g.n[2]);
    const long long q1 = q / g.n[2];
    const int j = (int)(q1 % g.n[1]);
    const int i = (int)(q1 / g.n[1]);
    const long long idx = (long long)(i + g.off[0]) * g.st[0] + (long long)(j + g.off[1]) * g.st[1] +
                          (long long)(k + g.off[2]) * g.st[2];
    double p[5];
#pragma unroll
    for (int v = 0; v < 5; ++v) p[v] = prims[idx + v * g.vst];
    r.add_cell(p, gamma, active_mask);
  }
  red_commit(r, red);
}

// stand-alone stage combination (time_integrator.py:108-227, RK3.py:49-60): whole buffer
// U <- a U + b U^n (stage > 0), then interior U += (dt m) rhs.  out may alias cons.
__global__ void __launch_bounds__(256) integrate_stage_kernel(const Geom g, const double* __restrict__ cons,
                                                              const double* __restrict__ cons_n,
                                                              const double* __restrict__ rhs, double* out, double ca,
                                                              double cb, int blend, double step) {
  const long long total = g.vst;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(q % g.ext[2]);
    const long long q1 = q / g.ext[2];
    const int j = (int)(q1 % g.ext[1]);
    const int i = (int)(q1 / g.ext[1]);
    const int ii = i - g.off[0], jj = j - g.off[1], kk = k - g.off[2];
    const bool interior = ii >= 0 && ii < g.n[0] && jj >= 0 && jj < g.n[1] && kk >= 0 && kk < g.n[2];
    const long long ridx = (long long)ii * g.rst[0] + (long long)jj * g.rst[1] + (long long)kk * g.rst[2];
#pragma unroll
    for (int v = 0; v < 5; ++v) {
      double u = cons[q + v * g.vst];
      if (blend) u = ca * u + cb * cons_n[q + v * g.vst];
      if (interior) u = u + step * rhs[ridx + v * g.rvst];
      out[q + v * g.vst] = u;
    }
  }
}

// volume forces for the stand-alone rhs entry points (jxf_compute_rhs): the stage path adds them in its epilogue
// from the stage's conservatives; here rho u is re-formed from the primitives (differs by rounding only)
__global__ void __launch_bounds__(256) gravity_rhs_kernel(const Geom g, const double* __restrict__ prims, double* rhs,
                                                          double g0, double g1, double g2) {
  const long long total = (long long)g.n[0] * g.n[1] * g.n[2];
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(q % g.n[2]);
    const long long q1 = q / g.n[2];
    const int j = (int)(q1 % g.n[1]);
    const int i = (int)(q1 / g.n[1]);
    const long long idx = (long long)(i + g.off[0]) * g.st[0] + (long long)(j + g.off[1]) * g.st[1] +
                          (long long)(k + g.off[2]) * g.st[2];
    const double rho = prims[idx];
    const double m0 = rho * prims[idx + g.vst], m1 = rho * prims[idx + 2 * g.vst], m2 = rho * prims[idx + 3 * g.vst];
    rhs[q + 1 * g.rvst] += g0 * rho;
    rhs[q + 2 * g.rvst] += g1 * rho;
    rhs[q + 3 * g.rvst] += g2 * rho;
    rhs[q + 4 * g.rvst] += (g0 * m0 + g1 * m1) + g2 * m2;
  }
}

// unfused stage update (no convective sweep to carry the epilogue: is_convective_flux = false): interior cells
// U <- a U + b U^n + (dt m) (rhs [+ gravity]), primitives recovered, reductions on the last stage; halos by jxf_halo_fill
struct UpdateArgs {
  const double* cons_in;
  const double* cons_n;
  const double* rhs;
  double* cons_out;
  double* prims_out;
  const double* dt;
  double* red;
  double ca, cb, dt_mult, gamma;
  double gravity[3];
  int blend, reduce, active_mask, volume_force;
};

__global__ void __launch_bounds__(256) update_stage_kernel(const Geom g, const UpdateArgs a) {
  const long long total = (long long)g.n[0] * g.n[1] * g.n[2];
  const double step = (*a.dt) * a.dt_mult;
  Red red;
  red.init();
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(q % g.n[2]);
    const long long q1 = q / g.n[2];
    const int j = (int)(q1 % g.n[1]);
    const int i = (int)(q1 / g.n[1]);
    const long long idx = (long long)(i + g.off[0]) * g.st[0] + (long long)(j + g.off[1]) * g.st[1] +
                          (long long)(k + g.off[2]) * g.st[2];
    double U0[5], U[5], p[5];
#pragma unroll
    for (int v = 0; v < 5; ++v) U0[v] = a.cons_in[idx + v * g.vst];
#pragma unroll
    for (int v = 0; v < 5; ++v) {
      double tot = a.rhs[q + v * g.rvst];
      if (a.volume_force) {
        if (v >= 1 && v <= 3) tot += a.gravity[v - 1] * U0[0];
        if (v == 4) tot += (a.gravity[0] * U0[1] + a.gravity[1] * U0[2]) + a.gravity[2] * U0[3];
      }
      double u = U0[v];
      if (a.blend) u = a.ca * u + a.cb * a.cons_n[idx + v * g.vst];
      U[v] = u + step * tot;
    }
    prims_from_cons(U, a.gamma, p);
#pragma unroll
    for (int v = 0; v < 5; ++v) {
      a.cons_out[idx + v * g.vst] = U[v];
      a.prims_out[idx + v * g.vst] = p[v];
    }
    if (a.reduce) red.add_cell(p, a.gamma, a.active_mask);
  }
  if (a.reduce) red_commit(red, a.red);
}

__global__ void reduce_reset_kernel(double* red) {
  red[0] = 0.0;
  red[1] = __longlong_as_double(0x7ff0000000000000LL);
  red[2] = red[1];
}

// time_step_size.py:103-109,154-155: dt = dx_min / (max + eps); dt *= CFL.  With the viscous / heat flux
// (:111-135): dt = min(dt, 3/14 dx^2 / (max nu + eps), 0.1 dx^2 / (max alpha + eps)), nu = mu / rho,
// alpha = lambda / (rho cp).  mu and lambda are constants on this path and x -> fl(mu / x), x -> fl(lambda /
// fl(x cp)) are monotone, so the maxima over the cells are attained at the minimum density red[1], bit for bit.
struct DtLimits {
  int visc, heat;
  double mu, lambda, cp;
};

__global__ void finish_step_kernel(double* red, double* dt, double* time, double* info, double dx_min, double cfl,
                                   double fixed_dt, DtLimits lim) {
  const double dt_used = *dt;
  if (time) *time += dt_used;
  if (info) {
    info[0] = red[0];
    info[1] = red[1];
    info[2] = red[2];
  }
  if (fixed_dt > 0.0) {
    *dt = fixed_dt;
  } else {
    double d = dx_min / (red[0] + kEps);
    const double dx2 = dx_min * dx_min;
    if (lim.visc) d = fmin(d, (3.0 / 14.0) * dx2 / (lim.mu / red[1] + kEps));
    if (lim.heat) d = fmin(d, 0.1 * dx2 / (lim.lambda / (red[1] * lim.cp) + kEps));
    d *= cfl;
    *dt = d;
  }
  red[0] = 0.0;
  red[1] = __longlong_as_double(0x7ff0000000000000LL);
  red[2] = red[1];
}

struct FaceArgs {
  double* prims;
  double* cons;
  double* slab;
  double gamma;
  int face;
  int unpack;
  int layers;             // layers exchanged (<= nh): the ones next to the face
  int lo1, n1, lo2, n2;   // transverse ranges in BUFFER coordinates: [lo, lo + n) along the two transverse axes
};

// slab layout (5, nh, n1, n2), layers in increasing buffer index along the face axis; the transverse ranges are
// the interior, optionally widened over the halo cells of a transverse axis (edge halos of the dissipative path)
__global__ void __launch_bounds__(128) face_slab_kernel(const Geom g, const FaceArgs a) {
  const int ax = a.face >> 1;
  const bool hi = (a.face & 1) == 0;
  const int t1 = (ax == 0) ? 1 : 0;
  const int t2 = (ax == 2) ? 1 : 2;
  const int n1 = a.n1, n2 = a.n2;
  const int nl = a.layers;
  const long long total = (long long)nl * n1 * n2;
  const int nh = g.nh, ext = g.ext[ax];
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total;
       q += (long long)gridDim.x * blockDim.x) {
    // slab order: layers slowest, except for faces of the CONTIGUOUS axis, whose nh layers are adjacent in memory:
    // there the layer index runs fastest so that a row's nh cells are one 8 nh-byte access (both ends of an
    // exchange run this kernel, so the order is private to it)
    int i1, i2, l;
    if (g.st[ax] == 1) {
      l = (int)(q % nl);
      const long long q1 = q / nl;
      i2 = (int)(q1 % n2);
      i1 = (int)(q1 / n2);
    } else {
      i2 = (int)(q % n2);
      const long long q1 = q / n2;
      i1 = (int)(q1 % n1);
      l = (int)(q1 / n1);
    }
    const long long tr = (long long)(i1 + a.lo1) * g.st[t1] + (long long)(i2 + a.lo2) * g.st[t2];
    if (!a.unpack) {
      const int src = hi ? (ext - nh - nl + l) : (nh + l);  // the nl interior layers adjacent to the face
      const long long is = tr + (long long)src * g.st[ax];
#pragma unroll
      for (int v = 0; v < 5; ++v) a.slab[q + v * total] = a.prims[is + v * g.vst];
    } else {
      const int dst = hi ? (ext - nh + l) : (nh - nl + l);  // the nl halo layers adjacent to the face
      const long long id = tr + (long long)dst * g.st[ax];
      double p[5], c[5];
#pragma unroll
      for (int v = 0; v < 5; ++v) p[v] = a.slab[q + v * total];
      cons_from_prims(p, a.gamma, c);
#pragma unroll
      for (int v = 0; v < 5; ++v) {
        a.prims[id + v * g.vst] = p[v];
        a.cons[id + v * g.vst] = c[v];
      }
    }
  }
}

__global__ void __launch_bounds__(256) fp64_probe_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0;
  double a4 = a0 + 4.0, a5 = a0 + 5.0, a6 = a0 + 6.0, a7 = a0 + 7.0;
  const double m = 0.9999999, c = 1e-7;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[blockIdx.x * (long long)blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

#ifndef JXF_REFERENCE_ORDER
// debug / test hook: accuracy of the MUFU seeds and of rcp_fast / rsqrt_fast: out = (n, 4)
__global__ void __launch_bounds__(128) math_debug_kernel(const double* __restrict__ x, long long n, double* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double a = x[i];
  double r, q;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(a));
  out[4 * i + 0] = r;
  out[4 * i + 1] = q;
  out[4 * i + 2] = rcp_fast(a);
  out[4 * i + 3] = rsqrt_fast(a);
}
#endif

// debug / test hook: the per-face device function on caller-supplied windows (n, 5, 6) -> (n, 5)
template <int A, int RECON, int RIEMANN>
__global__ void __launch_bounds__(128) face_flux_debug_kernel(const double* __restrict__ win, long long n, double gamma,
                                                              double* __restrict__ out, int opt) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double w[5][6], F[5];
#pragma unroll
  for (int v = 0; v < 5; ++v)
#pragma unroll
    for (int k = 0; k < 6; ++k) w[v][k] = win[(i * 5 + v) * 6 + k];
  const FluxLimArgs nofl = {nullptr, 0.0, 0.0};
  face_flux<A, RECON, RIEMANN>(w, gamma, F, opt, nofl);
#pragma unroll
  for (int v = 0; v < 5; ++v) out[i * 5 + v] = F[v];
}

}  // namespace jxf

#include "dissipative.cuh"

// ===========================================================================
// host side: plan + C ABI
// ===========================================================================
using namespace jxf;


static thread_local char g_err[512] = "";

int jxf_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

extern "C" const char* jxf_last_error(void) { return g_err; }
extern "C" int jxf_version(void) { return 110; }

extern "C" int jxf_create(const jxf_config* cfg, jxf_handle* out) {
  if (!cfg || !out) return fail(JXF_ERR_BAD_ARG, "jxf_create: null argument");
  if (cfg->nh < 3) return fail(JXF_ERR_BAD_ARG, "jxf_create: halo_cells=%d < 3 required by WENO5", cfg->nh);
  for (int i = 0; i < 3; ++i)
    if (cfg->n[i] < 1) return fail(JXF_ERR_BAD_ARG, "jxf_create: n[%d]=%d", i, cfg->n[i]);
  if (cfg->recon < JXF_RECON_PRIMITIVE || cfg->recon > JXF_RECON_CHAR_CONSERVATIVE)
    return fail(JXF_ERR_UNSUPPORTED, "jxf_create: reconstruction_variable id %d not implemented on the B200 path", cfg->recon);
  if (cfg->stencil < JXF_STENCIL_WENO5Z || cfg->stencil > JXF_STENCIL_TENO6A)
    return fail(JXF_ERR_UNSUPPORTED, "jxf_create: reconstruction_stencil id %d not implemented on the B200 path", cfg->stencil);
  if (cfg->riemann < JXF_RIEMANN_HLLC || cfg->riemann > JXF_RIEMANN_AUSMP)
    return fail(JXF_ERR_UNSUPPORTED, "jxf_create: riemann_solver id %d not implemented on the B200 path", cfg->riemann);
  if (cfg->signal_speed < JXF_SIGNAL_EINFELDT || cfg->signal_speed > JXF_SIGNAL_TORO)
    return fail(JXF_ERR_UNSUPPORTED, "jxf_create: signal_speed id %d not implemented on the B200 path", cfg->signal_speed);
  if (cfg->frozen_state != JXF_FROZEN_ARITHMETIC && cfg->frozen_state != JXF_FROZEN_ROE)
    return fail(JXF_ERR_UNSUPPORTED, "jxf_create: frozen_state id %d not implemented on the B200 path", cfg->frozen_state);
  if (cfg->convective_solver != JXF_SOLVER_GODUNOV && cfg->convective_solver != JXF_SOLVER_FLUX_SPLITTING)
    return fail(JXF_ERR_UNSUPPORTED, "jxf_create: convective_solver id %d not implemented on the B200 path", cfg->convective_solver);
  if (cfg->convective_solver == JXF_SOLVER_FLUX_SPLITTING) {
    if (cfg->flux_splitting < JXF_FS_ROE || cfg->flux_splitting > JXF_FS_LLF)
      return fail(JXF_ERR_UNSUPPORTED, "jxf_create: flux_splitting id %d not implemented on the B200 path", cfg->flux_splitting);
    if (cfg->flux_limiter != 0)
      return fail(JXF_ERR_UNSUPPORTED, "jxf_create: the positivity flux limiter is not implemented with FLUX-SPLITTING");
  }
  if (cfg->integrator < JXF_INT_EULER || cfg->integrator > JXF_INT_RK2_LS4)
    return fail(JXF_ERR_UNSUPPORTED, "jxf_create: integrator id %d not implemented on the B200 path", cfg->integrator);
  if (!(cfg->gamma > 1.0)) return fail(JXF_ERR_BAD_ARG, "jxf_create: gamma=%g", cfg->gamma);
  if (cfg->flux_limiter < JXF_FLUXLIM_NONE || cfg->flux_limiter > JXF_FLUXLIM_NASA)
    return fail(JXF_ERR_UNSUPPORTED, "jxf_create: flux_limiter id %d not implemented on the B200 path", cfg->flux_limiter);
  if (cfg->flux_partition < JXF_PARTITION_UNIFORM || cfg->flux_partition > JXF_PARTITION_CELLSIZE)
    return fail(JXF_ERR_UNSUPPORTED, "jxf_create: flux_partition id %d not implemented on the B200 path", cfg->flux_partition);
  if (cfg->no_convective_flux && !(cfg->viscous_flux || cfg->heat_flux))
    return fail(JXF_ERR_BAD_ARG, "jxf_create: no flux is active");
  if (cfg->viscous_flux || cfg->heat_flux) {
    if (!(cfg->gas_constant > 0.0)) return fail(JXF_ERR_BAD_ARG, "jxf_create: gas_constant=%g", cfg->gas_constant);
    if (cfg->dynamic_viscosity < 0.0 || cfg->thermal_conductivity < 0.0)
      return fail(JXF_ERR_BAD_ARG, "jxf_create: negative transport coefficient");
    if (cfg->nh < 4) return fail(JXF_ERR_BAD_ARG, "jxf_create: the dissipative fluxes need halo_cells >= 4 (2 + 2)");
  }
  jxf_solver* s = new (std::nothrow) jxf_solver;
  if (!s) return fail(JXF_ERR_BAD_ARG, "jxf_create: out of host memory");
  memset(s, 0, sizeof(*s));
  s->cfg = *cfg;
  Geom& g = s->g;
  g.nh = cfg->nh;
  s->n_active = 0;
  s->lane_axis = -1;
  for (int i = 0; i < 3; ++i) {
    g.n[i] = cfg->n[i];
    const bool act = cfg->n[i] > 1;
    g.ext[i] = act ? cfg->n[i] + 2 * cfg->nh : 1;
    g.off[i] = act ? cfg->nh : 0;
    if (act) {
      s->active[s->n_active++] = i;
      s->active_mask |= 1 << i;
      s->lane_axis = i;
      if (cfg->n[i] < cfg->nh)
        { delete s; return fail(JXF_ERR_BAD_ARG, "jxf_create: n[%d]=%d smaller than halo_cells", i, cfg->n[i]); }
    }
  }
  if (s->n_active == 0) { delete s; return fail(JXF_ERR_BAD_ARG, "jxf_create: no active axis"); }
  for (int f = 0; f < 6; ++f) {
    const int ax = f >> 1;
    const int b = cfg->bc[f];
    const bool act = cfg->n[ax] > 1;
    if (b < JXF_BC_INACTIVE || b > JXF_BC_DIRICHLET) { delete s; return fail(JXF_ERR_UNSUPPORTED, "jxf_create: boundary type id %d at face %d not implemented on the B200 path", b, f); }
    if (act && b == JXF_BC_INACTIVE) { delete s; return fail(JXF_ERR_BAD_ARG, "jxf_create: face %d of an active axis is INACTIVE", f); }
  }
  g.st[2] = 1;
  g.st[1] = g.ext[2];
  g.st[0] = (long long)g.ext[1] * g.ext[2];
  g.vst = (long long)g.ext[0] * g.ext[1] * g.ext[2];
  g.rst[2] = 1;
  g.rst[1] = g.n[2];
  g.rst[0] = (long long)g.n[1] * g.n[2];
  g.rvst = (long long)g.n[0] * g.n[1] * g.n[2];
  // RK tables: time_integration/euler.py, RK2.py:22-33, RK3.py:27-29, RK2_LS4.py:27-30
  if (cfg->integrator == JXF_INT_EULER) {
    s->stages = 1; s->dt_mult[0] = 1.0;
  } else if (cfg->integrator == JXF_INT_RK2_LS4) {
    // low-storage 4-stage scheme: every later stage restarts from U^n (blend (0, 1)): u^k = u^n + m_k dt L(u^{k-1})
    s->stages = 4; s->dt_mult[0] = 0.11; s->dt_mult[1] = 0.2766; s->dt_mult[2] = 0.5; s->dt_mult[3] = 1.0;
    for (int k = 1; k < 4; ++k) { s->blend[k][0] = 0.0; s->blend[k][1] = 1.0; }
  } else if (cfg->integrator == JXF_INT_RK2) {
    s->stages = 2; s->dt_mult[0] = 1.0; s->dt_mult[1] = 0.5;
    s->blend[1][0] = 0.5; s->blend[1][1] = 0.5;
  } else {
    s->stages = 3; s->dt_mult[0] = 1.0; s->dt_mult[1] = 0.25; s->dt_mult[2] = 2.0 / 3.0;
    s->blend[1][0] = 0.25; s->blend[1][1] = 0.75;
    s->blend[2][0] = 2.0 / 3.0; s->blend[2][1] = 1.0 / 3.0;
  }
  // Stage sweep order: the reference's x, y, z (space_solver.py:289-314), the fused epilogue on the last one.
  // JXF_SWEEP_ORDER=xzy (3-D) puts the epilogue on the marching y sweep instead and leaves the contiguous
  // z sweep a plain rhs accumulation; the sum (x + z) + y differs from (x + y) + z by rounding only.
  // Measured equal within noise at 512^3, so the reference order is the default.
  for (int k = 0; k < s->n_active; ++k) s->order[k] = s->active[k];
  {
    const char* so = getenv("JXF_SWEEP_ORDER");
    if (s->n_active == 3 && so && strcmp(so, "xzy") == 0) { s->order[0] = 0; s->order[1] = 2; s->order[2] = 1; }
  }
  s->tma_ok = !(getenv("JXF_NO_TMA") && atoi(getenv("JXF_NO_TMA")) != 0);
  s->no_march = false;
#ifdef JXF_WITH_STRIDED
  s->no_march = getenv("JXF_NO_MARCH") && atoi(getenv("JXF_NO_MARCH")) != 0;
#endif
  s->rows_group = getenv("JXF_ROWS_G") ? std::max(0, std::min(32, atoi(getenv("JXF_ROWS_G")))) : 0;
  s->no_lane_defer = !(getenv("JXF_LANE_DEFER") && atoi(getenv("JXF_LANE_DEFER")) != 0);   // opt-in: measured slower
  s->lean_images = getenv("JXF_LEAN_IMAGES") && atoi(getenv("JXF_LEAN_IMAGES")) != 0;
  s->no_tma_in = getenv("JXF_NO_TMA_IN") && atoi(getenv("JXF_NO_TMA_IN")) != 0;   // A/B: per-lane loads of the cell inputs
  s->no_plain = getenv("JXF_NO_PLAIN") && atoi(getenv("JXF_NO_PLAIN")) != 0;   // A/B: option-carrying instantiations
  s->force_rows = getenv("JXF_FORCE_ROWS") && atoi(getenv("JXF_FORCE_ROWS")) != 0;
  s->n_maps = 0;
  s->num_sms = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0) s->num_sms = sms;
  }
  // (The generic instantiations' stack frames are static -- no recursion -- and the driver sizes local memory per launch
  // from them, so the device-wide cudaLimitStackSize is left alone: raising it would reserve local memory for every
  // kernel of the process, torch's included.)
  (void)cudaGetLastError();
  *out = s;
  return JXF_OK;
}

static void prof_free(jxf_solver* s) {
  if (s->prof_start) {
    for (int i = 0; i < s->prof_cap; ++i) {
      cudaEventDestroy(s->prof_start[i]);
      cudaEventDestroy(s->prof_stop[i]);
    }
    delete[] s->prof_start;
    delete[] s->prof_stop;
    delete[] s->prof_kind;
    s->prof_start = s->prof_stop = nullptr;
    s->prof_kind = nullptr;
    s->prof_cap = s->prof_n = 0;
  }
}

extern "C" int jxf_destroy(jxf_handle h) {
  if (h) prof_free(h);
  delete h;
  return JXF_OK;
}

extern "C" int jxf_profile_enable(jxf_handle h, int enable) {
  if (!h) return fail(JXF_ERR_BAD_ARG, "jxf_profile_enable: null handle");
  if (enable && !h->prof_start) {
    const int cap = 4096;
    h->prof_start = new (std::nothrow) cudaEvent_t[cap];
    h->prof_stop = new (std::nothrow) cudaEvent_t[cap];
    h->prof_kind = new (std::nothrow) unsigned char[cap];
    if (!h->prof_start || !h->prof_stop || !h->prof_kind) return fail(JXF_ERR_BAD_ARG, "jxf_profile_enable: out of host memory");
    for (int i = 0; i < cap; ++i) {
      if (cudaEventCreate(&h->prof_start[i]) != cudaSuccess || cudaEventCreate(&h->prof_stop[i]) != cudaSuccess)
        return fail(JXF_ERR_CUDA, "jxf_profile_enable: cudaEventCreate failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    h->prof_cap = cap;
  }
  h->prof_on = enable ? 1 : 0;
  return JXF_OK;
}

extern "C" int jxf_profile_read(jxf_handle h, double* ms_sum, int64_t* timed, int64_t* launches, int reset) {
  if (!h) return fail(JXF_ERR_BAD_ARG, "jxf_profile_read: null handle");
  for (int k = 0; k < JXF_PROFILE_KINDS; ++k) {
    if (ms_sum) ms_sum[k] = 0.0;
    if (timed) timed[k] = 0;
    if (launches) launches[k] = h->launches[k];
  }
  for (int i = 0; i < h->prof_n; ++i) {
    if (cudaEventSynchronize(h->prof_stop[i]) != cudaSuccess)
      return fail(JXF_ERR_CUDA, "jxf_profile_read: %s", cudaGetErrorString(cudaGetLastError()));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->prof_start[i], h->prof_stop[i]);
    if (ms_sum) ms_sum[h->prof_kind[i]] += ms;
    if (timed) timed[h->prof_kind[i]]++;
  }
  if (reset) {
    h->prof_n = 0;
    for (int k = 0; k < JXF_PROFILE_KINDS; ++k) h->launches[k] = 0;
  }
  return JXF_OK;
}

extern "C" int64_t jxf_field_elems(jxf_handle h) { return h ? 5 * h->g.vst : -1; }
extern "C" int64_t jxf_rhs_elems(jxf_handle h) { return h ? 5 * h->g.rvst : -1; }
extern "C" int jxf_num_stages(jxf_handle h) { return h ? h->stages : -1; }

// ---------------------------------------------------------------------------
// TMA descriptors (driver entry point fetched at run time: the library does not link libcuda)
// ---------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
    (void)cudaGetLastError();
  }
  return fn;
}

// 4-D view (contiguous axis, faster transverse, slower transverse, variable) of a halo'd field buffer,
// box = kWinSlots cells x 1 x 1 x 5 variables.  Returns nullptr when TMA cannot describe the buffer
// (pitch not a multiple of 16 B, or no driver support) -- the cp.async loader is used instead.
const CUtensorMap* get_rows_map(jxf_solver* s, const double* base) {
  if (!s->tma_ok) return nullptr;
  for (int i = 0; i < s->n_maps; ++i)
    if (s->map_ptr[i] == base) return &s->map[i];
  PFN_encodeTiled enc = get_encode_fn();
  const Geom& g = s->g;
  if (!enc || (g.ext[2] * 8) % 16 != 0 || ((uintptr_t)base % 16) != 0) { s->tma_ok = false; return nullptr; }
  // physical layout is (5, ext0, ext1, ext2); an inactive trailing axis has extent 1, which keeps the
  // byte strides below valid; the contiguous ACTIVE axis may therefore be ext1 or ext0 with ext2 == 1.
  if (s->lane_axis != 2) { s->tma_ok = false; return nullptr; }    // 1-D / 2-D grids: cp.async loader
  cuuint64_t dims[4] = {(cuuint64_t)g.ext[2], (cuuint64_t)g.ext[1], (cuuint64_t)g.ext[0], 5};
  cuuint64_t strides[3] = {(cuuint64_t)g.ext[2] * 8, (cuuint64_t)g.ext[1] * g.ext[2] * 8, (cuuint64_t)g.vst * 8};
  cuuint32_t box[4] = {(cuuint32_t)kWinSlots, 1, 1, 5};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const int slot = s->n_maps < 8 ? s->n_maps : 7;
  CUresult rc = enc(&s->map[slot], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) { s->tma_ok = false; return nullptr; }
  s->map_ptr[slot] = base;
  if (s->n_maps < 8) s->n_maps++;
  return &s->map[slot];
}

// Tensor maps of the CELL INPUTS of the rows kernel's epilogue (sweep_kernels.cuh, RowsArgs::tma_in): a halo'd
// conservative buffer with a box of kInCells cells x 5 variables, or the interior-only (or slab-sized) rhs accumulator
// with a box of 32 cells x 5 variables.  Encoded per launch into the caller's storage (a few microseconds of host
// time); false when TMA cannot describe the buffer (odd extents: the per-lane loads are used instead).
bool encode_rows_input_map(const jxf_solver* s, CUtensorMap* out, const double* base, bool is_rhs, int rhs_planes,
                           long long rhs_vst) {
  PFN_encodeTiled enc = get_encode_fn();
  const Geom& g = s->g;
  if (!enc || !s->tma_ok || s->lane_axis != 2 || ((uintptr_t)base % 16) != 0) return false;
  cuuint64_t dims[4], strides[3];
  cuuint32_t box[4] = {(cuuint32_t)(is_rhs ? 32 : kInCells), 1, 1, 5};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  if (is_rhs) {
    if ((g.n[2] % 2) != 0 || (rhs_vst % 2) != 0) return false;
    dims[0] = g.n[2]; dims[1] = g.n[1]; dims[2] = rhs_planes; dims[3] = 5;
    strides[0] = (cuuint64_t)g.n[2] * 8; strides[1] = (cuuint64_t)g.n[1] * g.n[2] * 8; strides[2] = (cuuint64_t)rhs_vst * 8;
  } else {
    if ((g.ext[2] % 2) != 0 || (g.vst % 2) != 0) return false;
    dims[0] = g.ext[2]; dims[1] = g.ext[1]; dims[2] = g.ext[0]; dims[3] = 5;
    strides[0] = (cuuint64_t)g.ext[2] * 8; strides[1] = (cuuint64_t)g.ext[1] * g.ext[2] * 8; strides[2] = (cuuint64_t)g.vst * 8;
  }
  return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---------------------------------------------------------------------------
// sweep dispatch
// ---------------------------------------------------------------------------
// launch_sweep<A, RECON, RIEMANN, EPI> is instantiated in sweep_inst.cu, one translation unit per (A, RECON)
#ifndef JXF_RECON_LIMIT
#define JXF_RECON_LIMIT 6     // RECON templates 0 .. JXF_RECON_LIMIT - 1 are in this build
#endif
#define JXF_EXTERN_SWEEP(A, R, S, E) extern template int launch_sweep<A, R, S, E>(const jxf_solver*, SweepArgs, cudaStream_t);
#ifdef JXF_TUNE_ONLY
JXF_SWEEPS_TUNE(JXF_EXTERN_SWEEP, 0) JXF_SWEEPS_TUNE(JXF_EXTERN_SWEEP, 1) JXF_SWEEPS_TUNE(JXF_EXTERN_SWEEP, 2)
#else
#if JXF_RECON_LIMIT < 6     // checking builds with the WENO5-Z instantiations only (build.py build_reforder)
#define JXF_EXTERN_AXIS(A)                                                                                   \
  JXF_SWEEPS_OF(JXF_EXTERN_SWEEP, A, 0) JXF_SWEEPS_OF(JXF_EXTERN_SWEEP, A, 1)                                 \
  JXF_SWEEPS_PLAIN_OF(JXF_EXTERN_SWEEP, A, 0) JXF_SWEEPS_PLAIN_OF(JXF_EXTERN_SWEEP, A, 1)
#else
#define JXF_EXTERN_AXIS(A)                                                                                   \
  JXF_SWEEPS_OF(JXF_EXTERN_SWEEP, A, 0) JXF_SWEEPS_OF(JXF_EXTERN_SWEEP, A, 1) JXF_SWEEPS_OF(JXF_EXTERN_SWEEP, A, 2) \
  JXF_SWEEPS_OF(JXF_EXTERN_SWEEP, A, 3) JXF_SWEEPS_OF(JXF_EXTERN_SWEEP, A, 4) JXF_SWEEPS_OF(JXF_EXTERN_SWEEP, A, 5) \
  JXF_SWEEPS_PLAIN_OF(JXF_EXTERN_SWEEP, A, 0) JXF_SWEEPS_PLAIN_OF(JXF_EXTERN_SWEEP, A, 1)                     \
  JXF_SWEEPS_PLAIN_OF(JXF_EXTERN_SWEEP, A, 2) JXF_SWEEPS_PLAIN_OF(JXF_EXTERN_SWEEP, A, 3)
#endif
JXF_EXTERN_AXIS(0) JXF_EXTERN_AXIS(1) JXF_EXTERN_AXIS(2)
#undef JXF_EXTERN_AXIS
#endif
#undef JXF_EXTERN_SWEEP

// stencil id in the option word: bits 11-14, the fifth id bit at bit 22 (numerics.cuh stencil_id)
static int stencil_bits(int stencil) { return ((stencil & 15) << 11) | ((stencil >> 4) << 22); }

// Godunov setups the tuned instantiations do not cover (numerics.cuh STENCIL_GENERIC)
static bool generic_path(const jxf_solver* s) {
  return s->cfg.stencil >= JXF_STENCIL_WENO1 || s->cfg.recon >= JXF_RECON_CONSERVATIVE || s->cfg.frozen_state == JXF_FROZEN_ROE;
}

template <int A, int RECON, int RIEMANN>
static int dispatch_epi(const jxf_solver* s, const SweepArgs& a, int epi, cudaStream_t st) {
  return epi ? launch_sweep<A, RECON, RIEMANN, 1>(s, a, st) : launch_sweep<A, RECON, RIEMANN, 0>(s, a, st);
}
// HLLC + EINFELDT with option word 0 (no limiter, no alternative signal speed) on a tuned stencil: the RIEMANN_HLLC_PLAIN
// instantiations (numerics.cuh), and for the usual epilogue (earlier axes' sum present, no volume force) the
// instantiations with compile-time blend / reduce flags (sweep_kernels.cuh EpiFlags).  Same arithmetic.
template <int A, int RECON>
static int dispatch_plain(const jxf_solver* s, const SweepArgs& a, int epi, cudaStream_t st) {
  if (!epi) return launch_sweep<A, RECON, RIEMANN_HLLC_PLAIN, 0>(s, a, st);
  if (!a.has_prev || a.volume_force) return launch_sweep<A, RECON, RIEMANN_HLLC_PLAIN, 1>(s, a, st);
  switch ((a.blend ? 1 : 0) | (a.reduce ? 2 : 0)) {
    case 0: return launch_sweep<A, RECON, RIEMANN_HLLC_PLAIN, 2>(s, a, st);
    case 1: return launch_sweep<A, RECON, RIEMANN_HLLC_PLAIN, 3>(s, a, st);
    case 2: return launch_sweep<A, RECON, RIEMANN_HLLC_PLAIN, 4>(s, a, st);
    default: return launch_sweep<A, RECON, RIEMANN_HLLC_PLAIN, 5>(s, a, st);
  }
}
// The kernel instantiation a configuration runs in (also reported by jxf_debug_dispatch, so that the host simulation
// of the device functions, tests/hostsim, can be checked to use the same template parameters and option word):
//   RECON   0..3 = reconstruction variable + 2 * stencil for the tuned WENO5-Z / WENO5-JS forms; 4 / 5 = the generic
//           instantiations (every other stencil, the conservative variables, the ROE frozen state, FLUX-SPLITTING);
//   RIEMANN HLLC, or RUSANOV for everything else (Rusanov, HLL, HLLC-LM, AUSM+ as run-time variants; FLUX-SPLITTING).
static int recon_template_of(const jxf_solver* s) {
  if (s->cfg.convective_solver == JXF_SOLVER_FLUX_SPLITTING) return 4;
  if (generic_path(s)) return 4 + (s->cfg.recon & 1);
  return s->cfg.recon + 2 * s->cfg.stencil;
}
static int riemann_template_of(const jxf_solver* s) {
  if (s->cfg.convective_solver == JXF_SOLVER_FLUX_SPLITTING) return RIEMANN_RUSANOV;
  return s->cfg.riemann == JXF_RIEMANN_HLLC ? RIEMANN_HLLC : RIEMANN_RUSANOV;
}

template <int A, int RECON>
static int dispatch_riemann(const jxf_solver* s, const SweepArgs& a, int epi, cudaStream_t st) {
  if constexpr (RECON < 4) {
    if (riemann_template_of(s) == RIEMANN_HLLC && a.limiter == 0 && !s->no_plain) return dispatch_plain<A, RECON>(s, a, epi, st);
  }
#ifdef JXF_TUNE_ONLY   // tuning builds instantiate the bench variant only (CHAR-PRIMITIVE + HLLC)
  if (s->cfg.riemann != JXF_RIEMANN_HLLC) return fail(JXF_ERR_UNSUPPORTED, "tuning build: HLLC only");
  return dispatch_epi<A, RECON, RIEMANN_HLLC>(s, a, epi, st);
#else
  return riemann_template_of(s) == RIEMANN_HLLC ? dispatch_epi<A, RECON, RIEMANN_HLLC>(s, a, epi, st)
                                                : dispatch_epi<A, RECON, RIEMANN_RUSANOV>(s, a, epi, st);
#endif
}
template <int A>
static int dispatch_recon(const jxf_solver* s, const SweepArgs& a, int epi, cudaStream_t st) {
  // kernels' RECON parameter = reconstruction variable + 2 * stencil (numerics.cuh)
#ifdef JXF_TUNE_ONLY
  if (s->cfg.recon != JXF_RECON_CHAR_PRIMITIVE || s->cfg.stencil != JXF_STENCIL_WENO5Z)
    return fail(JXF_ERR_UNSUPPORTED, "tuning build: WENO5-Z CHAR-PRIMITIVE only");
  return dispatch_riemann<A, RECON_CHAR_PRIMITIVE>(s, a, epi, st);
#else
  // FLUX-SPLITTING is a run-time branch of the generic instantiations' face flux (numerics.cuh flux_splitting_flux);
  // every stencil other than the two tuned WENO5 forms, the conservative reconstruction variables and the ROE frozen
  // state run in the STENCIL_GENERIC instantiations (RECON 4 / 5), selected at run time by the option word (base_args)
  if (recon_template_of(s) >= JXF_RECON_LIMIT)
    return fail(JXF_ERR_UNSUPPORTED, "this (checking) build holds the RECON < %d kernel instantiations only", JXF_RECON_LIMIT);
  switch (recon_template_of(s)) {
    case 0: return dispatch_riemann<A, 0>(s, a, epi, st);
    case 1: return dispatch_riemann<A, 1>(s, a, epi, st);
#if JXF_RECON_LIMIT >= 6
    case 2: return dispatch_riemann<A, 2>(s, a, epi, st);
    case 3: return dispatch_riemann<A, 3>(s, a, epi, st);
    case 4: return dispatch_riemann<A, 4>(s, a, epi, st);
    default: return dispatch_riemann<A, 5>(s, a, epi, st);
#else
    default: return JXF_ERR_UNSUPPORTED;
#endif
  }
#endif
}
static int dispatch_axis(const jxf_solver* s, int axis, const SweepArgs& a, int epi, cudaStream_t st) {
  switch (axis) {
    case 0: return dispatch_recon<0>(s, a, epi, st);
    case 1: return dispatch_recon<1>(s, a, epi, st);
    default: return dispatch_recon<2>(s, a, epi, st);
  }
}

static SweepArgs base_args(const jxf_solver* s, int axis, const double* prims, double* rhs) {
  SweepArgs a;
  memset(&a, 0, sizeof(a));
  a.prims = prims;
  a.rhs = rhs;
  a.gamma = s->cfg.gamma;
  a.inv_dx = s->cfg.inv_dx[axis];
  a.active_mask = s->active_mask;
  // packed face-flux options (numerics.cuh face_flux `opt`): limiter mode | signal speed << 4
  a.limiter = (s->cfg.interpolation_limiter ? (s->cfg.limit_velocity ? 2 : 1) : 0) | (s->cfg.signal_speed << 4) |
              ((s->cfg.riemann == JXF_RIEMANN_HLL ? 1 : 0) << 8) |     // HLL rides on the RUSANOV instantiations
              ((s->cfg.riemann == JXF_RIEMANN_HLLCLM ? RIEMANN_ALT_HLLCLM                 // ... and so do HLLC-LM, AUSM+
                : s->cfg.riemann == JXF_RIEMANN_AUSMP ? RIEMANN_ALT_AUSMP : 0) << 15) |
              (s->cfg.flux_limiter << 9) |
              // generic instantiations: stencil id (numerics.cuh ALT_*) | reconstruction variable << 19 | ROE << 21
              (generic_path(s) ? stencil_bits(s->cfg.stencil) | (s->cfg.recon << 19) | (s->cfg.frozen_state << 21) : 0);
  if (s->cfg.convective_solver == JXF_SOLVER_FLUX_SPLITTING)       // stencil id | eigenvalue choice << 17 | ROE << 21
    a.limiter = stencil_bits(s->cfg.stencil) | (s->cfg.flux_splitting << 17) | (s->cfg.frozen_state << 21);
  // positivity flux limiter: lambda = dt / dx * sigma (limiter_flux.py:202-205, compute_partition :681-720)
  a.fl.dt = s->dt_bound;
  a.fl.inv_dx = s->cfg.inv_dx[axis];
  a.fl.sigma = (double)s->n_active;
  if (s->cfg.flux_partition == JXF_PARTITION_CELLSIZE) {
    double sum = 0.0;
    for (int k = 0; k < s->n_active; ++k) sum += s->cfg.inv_dx[s->active[k]];
    a.fl.sigma = sum / s->cfg.inv_dx[axis];
  }
  return a;
}

// ---------------------------------------------------------------------------
// dissipative sweeps (dissipative.cuh)
// ---------------------------------------------------------------------------
static bool dissipative(const jxf_solver* s) { return s->cfg.viscous_flux || s->cfg.heat_flux; }

template <int A>
static int launch_dissipative(const jxf_solver* s, const double* prims, double* rhs, int accumulate, cudaStream_t st) {
  const Geom& g = s->g;
  const long long h0 = g.off[0] * g.st[0] + g.off[1] * g.st[1] + g.off[2] * g.st[2];
  SweepGeom sg;
  memset(&sg, 0, sizeof(sg));
  sg.axA = A; sg.nA = g.n[A]; sg.sA = g.st[A]; sg.rA = g.rst[A];
  sg.vst = g.vst; sg.rvst = g.rvst;
  ViscArgs a;
  memset(&a, 0, sizeof(a));
  a.prims = prims + h0;
  a.rhs = rhs;
  a.mu1 = s->cfg.dynamic_viscosity;
  a.mu2 = s->cfg.bulk_viscosity - 2.0 / 3.0 * s->cfg.dynamic_viscosity;     // source_term_solver.py:524
  a.lambda = s->cfg.thermal_conductivity;
  a.gas_constant = s->cfg.gas_constant;
  a.visc = s->cfg.viscous_flux;
  a.heat = s->cfg.heat_flux;
  a.heat_prod = s->cfg.viscous_heat_production;
  a.active_mask = s->active_mask;
  a.accumulate = accumulate ? 1 : 0;
  a.inv_dxA = s->cfg.inv_dx[A];
  ProfScope prof(s, JXF_PROFILE_DISSIPATIVE, st);
  if (A != s->lane_axis) {
    const int C = s->lane_axis, O = 3 - A - C;
    sg.ax1 = O; sg.n1 = g.n[O]; sg.s1 = g.st[O]; sg.r1 = g.rst[O];
    sg.ax2 = C; sg.n2 = g.n[C]; sg.s2 = g.st[C]; sg.r2 = g.rst[C];
    a.inv_dx1 = s->cfg.inv_dx[O];
    a.inv_dx2 = s->cfg.inv_dx[C];
    const int bx = (sg.n2 + 31) / 32, by = (sg.n1 + 3) / 4;
    if (by > 65535) return fail(JXF_ERR_UNSUPPORTED, "dissipative sweep: transverse extent too large");
    // enough CTAs for ~8 per SM; every chunk re-reads a 3-cell prologue
    const long long tiles = (long long)bx * by;
    int chunks = (int)std::min<long long>(std::max<long long>(1, (8LL * s->num_sms + tiles - 1) / tiles),
                                          std::max(1, std::min(g.n[A] / 16, 65535)));
    a.chunk_len = (g.n[A] + chunks - 1) / chunks;
    chunks = (g.n[A] + a.chunk_len - 1) / a.chunk_len;
    visc_march<A><<<dim3(bx, by, chunks), dim3(32, 4), 0, st>>>(sg, a);
  } else {
    const int T1 = (A == 0) ? 1 : 0, T2 = (A == 2) ? 1 : 2;
    sg.ax1 = T1; sg.n1 = g.n[T1]; sg.s1 = g.st[T1]; sg.r1 = g.rst[T1];
    sg.ax2 = T2; sg.n2 = g.n[T2]; sg.s2 = g.st[T2]; sg.r2 = g.rst[T2];
    a.inv_dx1 = s->cfg.inv_dx[T1];
    a.inv_dx2 = s->cfg.inv_dx[T2];
    // CTA = R consecutive rows x one segment of <= 252 cells (+4 stencil cells)
    int seg_max = 252;
    if (getenv("JXF_VISC_SEG")) seg_max = std::max(28, std::min(252, atoi(getenv("JXF_VISC_SEG"))));
    const int nseg = (g.n[A] + seg_max - 1) / seg_max;
    a.seg_len = (g.n[A] + nseg - 1) / nseg;
    const int threads = ((a.seg_len + 4 + 31) / 32) * 32;
    int R = 1;      // measured at 512^3: 1 row per CTA 6.7 ms, 2 rows 8.6 ms, 4 rows 10.1 ms (more CTAs overlap the phases)
    if (getenv("JXF_VISC_ROWS_R")) R = std::max(1, std::min(256 / threads, atoi(getenv("JXF_VISC_ROWS_R"))));
    const long long rows = (long long)sg.n1 * sg.n2;
    const long long groups = (rows + R - 1) / R;
    if (groups > 0x7fffffffLL || nseg > 65535) return fail(JXF_ERR_UNSUPPORTED, "dissipative sweep: grid too large");
    const size_t smem = (size_t)R * threads * 14 * sizeof(double);     // <= 28 KB
    static bool attr_set = false;                                       // per instantiation (per axis)
    if (!attr_set) {
      if (cudaFuncSetAttribute(visc_rows<A>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 14 * (int)sizeof(double)) != cudaSuccess)
        return fail(JXF_ERR_CUDA, "dissipative sweep: cudaFuncSetAttribute: %s", cudaGetErrorString(cudaGetLastError()));
      attr_set = true;
    }
    visc_rows<A><<<dim3((unsigned)groups, nseg), dim3(threads, R), smem, st>>>(sg, a);
  }
  return check_launch("dissipative sweep");
}

static int dissipative_sweep(const jxf_solver* s, int axis, const double* prims, double* rhs, int accumulate, cudaStream_t st) {
  switch (axis) {
    case 0: return launch_dissipative<0>(s, prims, rhs, accumulate, st);
    case 1: return launch_dissipative<1>(s, prims, rhs, accumulate, st);
    default: return launch_dissipative<2>(s, prims, rhs, accumulate, st);
  }
}

extern "C" int jxf_dissipative_sweep(jxf_handle h, int axis, const double* prims, double* rhs, int accumulate, void* stream) {
  if (!h || !prims || !rhs) return fail(JXF_ERR_BAD_ARG, "jxf_dissipative_sweep: null argument");
  if (axis < 0 || axis > 2 || h->g.n[axis] <= 1) return fail(JXF_ERR_BAD_ARG, "jxf_dissipative_sweep: axis %d is not active", axis);
  if (!dissipative(h)) return fail(JXF_ERR_BAD_ARG, "jxf_dissipative_sweep: neither viscous nor heat flux is configured");
  return dissipative_sweep(h, axis, prims, rhs, accumulate, (cudaStream_t)stream);
}

extern "C" int jxf_halo_fill_edges(jxf_handle h, double* prims, double* cons, void* stream) {
  if (!h || !prims || !cons) return fail(JXF_ERR_BAD_ARG, "jxf_halo_fill_edges: null argument");
  if (h->n_active < 2) return JXF_OK;
  EdgeArgs a;
  a.prims = prims;
  a.cons = cons;
  a.gamma = h->cfg.gamma;
  int nmax = 1;
  for (int f = 0; f < 6; ++f) a.bc[f] = (h->g.n[f >> 1] > 1) ? h->cfg.bc[f] : JXF_BC_INACTIVE;
  for (int i = 0; i < 3; ++i) nmax = std::max(nmax, h->g.n[i]);
  const long long cells = (long long)h->g.nh * h->g.nh * nmax;
  const int bx = (int)std::min<long long>((cells + 127) / 128, 148 * 4);
  ProfScope prof(h, JXF_PROFILE_HALO, (cudaStream_t)stream);
  halo_fill_edges_kernel<<<dim3(bx, 12), 128, 0, (cudaStream_t)stream>>>(h->g, a);
  return check_launch("halo_fill_edges");
}

extern "C" int jxf_temperature(jxf_handle h, const double* prims, double* temperature, void* stream) {
  if (!h || !prims || !temperature) return fail(JXF_ERR_BAD_ARG, "jxf_temperature: null argument");
  if (!(h->cfg.gas_constant > 0.0)) return fail(JXF_ERR_BAD_ARG, "jxf_temperature: gas_constant not configured");
  const int bx = (int)std::min<long long>((h->g.vst + 255) / 256, 148 * 8);
  ProfScope prof(h, JXF_PROFILE_OTHER, (cudaStream_t)stream);
  temperature_kernel<<<bx, 256, 0, (cudaStream_t)stream>>>(prims, temperature, h->g.vst, h->cfg.gas_constant);
  return check_launch("temperature");
}

extern "C" int jxf_sweep(jxf_handle h, int axis, const double* prims, double* rhs, int accumulate, void* stream) {
  if (!h || !prims || !rhs) return fail(JXF_ERR_BAD_ARG, "jxf_sweep: null argument");
  if (axis < 0 || axis > 2 || h->g.n[axis] <= 1) return fail(JXF_ERR_BAD_ARG, "jxf_sweep: axis %d is not active", axis);
  if (h->cfg.no_convective_flux)      // flux_xi = 0 + dissipative part only (space_solver.py:517-584)
    return dissipative_sweep(h, axis, prims, rhs, accumulate, (cudaStream_t)stream);
  if (h->cfg.flux_limiter && !h->dt_bound)
    return fail(JXF_ERR_BAD_ARG, "jxf_sweep: the flux limiter needs the time step (jxf_bind_timestep)");
  SweepArgs a = base_args(h, axis, prims, rhs);
  a.accumulate = accumulate ? 1 : 0;
  int rc = dispatch_axis(h, axis, a, 0, (cudaStream_t)stream);
  // compute_rhs_xi folds the axis' viscous / heat flux into the same divergence (space_solver.py:567-599)
  if (rc == JXF_OK && dissipative(h)) rc = dissipative_sweep(h, axis, prims, rhs, 1, (cudaStream_t)stream);
  return rc;
}

extern "C" int jxf_sweep_range(jxf_handle h, int axis, int lo, int hi, const double* prims, double* rhs, int accumulate,
                               void* stream) {
  if (!h || !prims || !rhs) return fail(JXF_ERR_BAD_ARG, "jxf_sweep_range: null argument");
  if (axis < 0 || axis > 2 || h->g.n[axis] <= 1) return fail(JXF_ERR_BAD_ARG, "jxf_sweep_range: axis %d is not active", axis);
  if (lo < 0 || hi > h->g.n[axis] || lo >= hi) return fail(JXF_ERR_BAD_ARG, "jxf_sweep_range: bad range [%d, %d)", lo, hi);
  if (dissipative(h)) return fail(JXF_ERR_UNSUPPORTED, "jxf_sweep_range: not available with the viscous / heat flux");
  if (axis == h->lane_axis) {
    if (lo != 0 || hi != h->g.n[axis])
      return fail(JXF_ERR_UNSUPPORTED, "jxf_sweep_range: partial ranges are not supported along the contiguous axis");
    return jxf_sweep(h, axis, prims, rhs, accumulate, stream);
  }
  if (h->cfg.flux_limiter && !h->dt_bound)
    return fail(JXF_ERR_BAD_ARG, "jxf_sweep_range: the flux limiter needs the time step (jxf_bind_timestep)");
  SweepArgs a = base_args(h, axis, prims, rhs);
  a.accumulate = accumulate ? 1 : 0;
  a.range_lo = lo;
  a.range_hi = hi;
  return dispatch_axis(h, axis, a, 0, (cudaStream_t)stream);
}

extern "C" int jxf_bind_timestep(jxf_handle h, const double* dt) {
  if (!h) return fail(JXF_ERR_BAD_ARG, "jxf_bind_timestep: null handle");
  h->dt_bound = dt;
  return JXF_OK;
}

extern "C" int jxf_compute_rhs(jxf_handle h, const double* prims, double* rhs, void* stream) {
  if (!h || !prims || !rhs) return fail(JXF_ERR_BAD_ARG, "jxf_compute_rhs: null argument");
  for (int k = 0; k < h->n_active; ++k) {
    int rc = jxf_sweep(h, h->active[k], prims, rhs, k > 0, stream);
    if (rc) return rc;
  }
  if (h->cfg.volume_force) {
    const int bx = (int)std::min<long long>((h->g.rvst + 255) / 256, 148 * 8);
    ProfScope prof(h, JXF_PROFILE_OTHER, (cudaStream_t)stream);
    gravity_rhs_kernel<<<bx, 256, 0, (cudaStream_t)stream>>>(h->g, prims, rhs, h->cfg.gravity[0], h->cfg.gravity[1],
                                                            h->cfg.gravity[2]);
    return check_launch("gravity_rhs");
  }
  return JXF_OK;
}

// Per-face boundary data (sweep_kernels.cuh FaceData): the caller owns the device arrays and keeps them alive while the
// handle uses them.  ops = 0 (or data = null) clears the face.
extern "C" int jxf_set_face_data(jxf_handle h, int face, int ops, const double* data_dev, const unsigned char* mask_dev) {
  if (!h || face < 0 || face > 5) return fail(JXF_ERR_BAD_ARG, "jxf_set_face_data: bad argument");
  if (ops < 0 || ops >= (1 << 10)) return fail(JXF_ERR_BAD_ARG, "jxf_set_face_data: ops=%d", ops);
  for (int v = 0; v < 5; ++v)
    if (((ops >> (2 * v)) & 3) == 3) return fail(JXF_ERR_BAD_ARG, "jxf_set_face_data: op 3 of variable %d is undefined", v);
  if (ops && !data_dev) return fail(JXF_ERR_BAD_ARG, "jxf_set_face_data: ops without data");
  const int b = h->cfg.bc[face];
  if (ops && (b == JXF_BC_PERIODIC || b == JXF_BC_NEIGHBOR || b == JXF_BC_INACTIVE))
    return fail(JXF_ERR_BAD_ARG, "jxf_set_face_data: face %d carries no physical boundary rule to build on", face);
  h->face_data.ops[face] = data_dev ? ops : 0;
  h->face_data.data[face] = ops ? data_dev : nullptr;
  h->face_data.mask[face] = ops ? mask_dev : nullptr;
  h->has_face_data = 0;
  for (int f = 0; f < 6; ++f) h->has_face_data |= (h->face_data.ops[f] != 0);
  return JXF_OK;
}

// ---------------------------------------------------------------------------
// Peer-memory halo exchange (multi-GPU, one process per GPU, neighbours' buffers mapped with CUDA IPC): the fused
// epilogue stores the halo images of NEIGHBOR faces straight into the neighbour's output buffers over NVLink
// (halo_images_axis); what remains of the "exchange" is a pair of flags per shared face.
//   jxf_peer_signal: after a stage's kernels (stream order), publish `epoch` in every neighbour's flag slot.
//   jxf_peer_wait:   before the first kernel that reads the halos, spin until every neighbour has published `epoch`.
// Every rank signals stage k before it waits for stage k, so the waits cannot form a cycle.
// Replaces halos/inner/material.py:30-93 (ppermute of the face slabs) without staging buffers or a collective.
// ---------------------------------------------------------------------------
struct PeerFlags {
  long long* slot[6];      // the neighbour's flag word this block writes, per face (null: none)
};

__global__ void peer_signal_kernel(PeerFlags pf, long long epoch) {
  const int f = threadIdx.x;
  if (f < 6 && pf.slot[f]) {
    __threadfence_system();                     // this GPU's earlier stores (incl. the remote halo images) first
    *reinterpret_cast<volatile long long*>(pf.slot[f]) = epoch;
  }
}

__global__ void peer_wait_kernel(const long long* flags, int face_mask, long long epoch) {
  const int f = threadIdx.x;
  if (f < 6 && ((face_mask >> f) & 1)) {
    const volatile long long* w = flags + f;
    while (*w < epoch) __nanosleep(100);
  }
  __threadfence_system();
}

// Mapping the neighbours' buffers: jxf_peer_export names the device allocation that holds `ptr` (a 64-byte CUDA IPC
// handle + the byte offset of ptr inside it); jxf_peer_import opens such a handle IN THE CONTEXT OF THE CURRENT DEVICE
// (cudaIpcMemLazyEnablePeerAccess: the driver enables NVLink peer access to the owner's device) and returns the mapped
// address of the same byte.  One allocation is opened once per process (handles are cached), jxf_peer_release unmaps all.
namespace {
struct PeerMapping { unsigned char handle[64]; void* base; int device; };
std::mutex g_peer_mu;
std::vector<PeerMapping> g_peer_maps;
typedef CUresult (*PFN_getAddressRange)(CUdeviceptr*, size_t*, CUdeviceptr);
}

extern "C" int jxf_peer_export(const void* ptr, unsigned char* handle_out, int64_t* offset_out) {
  if (!ptr || !handle_out || !offset_out) return fail(JXF_ERR_BAD_ARG, "jxf_peer_export: null argument");
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fp, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess || !fp) {
    (void)cudaGetLastError();
    return fail(JXF_ERR_CUDA, "jxf_peer_export: cuMemGetAddressRange is not available");
  }
  CUdeviceptr base = 0;
  size_t size = 0;
  CUresult r = ((PFN_getAddressRange)fp)(&base, &size, (CUdeviceptr)(uintptr_t)ptr);
  if (r != CUDA_SUCCESS) return fail(JXF_ERR_CUDA, "jxf_peer_export: cuMemGetAddressRange failed (%d)", (int)r);
  cudaIpcMemHandle_t hd;
  cudaError_t e = cudaIpcGetMemHandle(&hd, (void*)(uintptr_t)base);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return fail(JXF_ERR_CUDA, "jxf_peer_export: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
  }
  static_assert(sizeof(hd) == 64, "CUDA IPC handles are 64 bytes");
  memcpy(handle_out, &hd, 64);
  *offset_out = (int64_t)((uintptr_t)ptr - (uintptr_t)base);
  return JXF_OK;
}

extern "C" int jxf_peer_import(const unsigned char* handle, int64_t offset, void** ptr_out) {
  if (!handle || !ptr_out || offset < 0) return fail(JXF_ERR_BAD_ARG, "jxf_peer_import: bad argument");
  int dev = -1;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(g_peer_mu);
  for (const PeerMapping& m : g_peer_maps)
    if (m.device == dev && memcmp(m.handle, handle, 64) == 0) {
      *ptr_out = (char*)m.base + offset;
      return JXF_OK;
    }
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle, 64);
  void* base = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&base, hd, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return fail(JXF_ERR_CUDA, "jxf_peer_import: cudaIpcOpenMemHandle on device %d: %s", dev, cudaGetErrorString(e));
  }
  PeerMapping m;
  memcpy(m.handle, handle, 64);
  m.base = base;
  m.device = dev;
  g_peer_maps.push_back(m);
  *ptr_out = (char*)base + offset;
  return JXF_OK;
}

extern "C" int jxf_peer_release(void) {
  std::lock_guard<std::mutex> lock(g_peer_mu);
  int rc = JXF_OK;
  for (const PeerMapping& m : g_peer_maps) {
    int cur = -1;
    cudaGetDevice(&cur);
    if (cur != m.device) cudaSetDevice(m.device);
    if (cudaIpcCloseMemHandle(m.base) != cudaSuccess) {
      (void)cudaGetLastError();
      rc = fail(JXF_ERR_CUDA, "jxf_peer_release: cudaIpcCloseMemHandle failed");
    }
    if (cur != m.device) cudaSetDevice(cur);
  }
  g_peer_maps.clear();
  return rc;
}

extern "C" int jxf_set_peer_halo(jxf_handle h, int face, double* peer_prims_out, double* peer_cons_out) {
  if (!h || face < 0 || face > 5) return fail(JXF_ERR_BAD_ARG, "jxf_set_peer_halo: bad argument");
  if ((peer_prims_out == nullptr) != (peer_cons_out == nullptr))
    return fail(JXF_ERR_BAD_ARG, "jxf_set_peer_halo: both buffers or none");
  if (peer_prims_out && h->cfg.bc[face] != JXF_BC_NEIGHBOR)
    return fail(JXF_ERR_BAD_ARG, "jxf_set_peer_halo: face %d is not shared with another block", face);
  h->peer_prims[face] = peer_prims_out;
  h->peer_cons[face] = peer_cons_out;
  return JXF_OK;
}

extern "C" int jxf_peer_signal(jxf_handle h, int64_t* const* neighbour_flag_slots, int64_t epoch, void* stream) {
  if (!h || !neighbour_flag_slots) return fail(JXF_ERR_BAD_ARG, "jxf_peer_signal: null argument");
  PeerFlags pf;
  for (int f = 0; f < 6; ++f) pf.slot[f] = reinterpret_cast<long long*>(neighbour_flag_slots[f]);
  ProfScope prof(h, JXF_PROFILE_OTHER, (cudaStream_t)stream);
  peer_signal_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pf, (long long)epoch);
  return check_launch("peer_signal");
}

extern "C" int jxf_peer_wait(jxf_handle h, const int64_t* flags, int face_mask, int64_t epoch, void* stream) {
  if (!h || !flags) return fail(JXF_ERR_BAD_ARG, "jxf_peer_wait: null argument");
  ProfScope prof(h, JXF_PROFILE_OTHER, (cudaStream_t)stream);
  peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<const long long*>(flags), face_mask, (long long)epoch);
  return check_launch("peer_wait");
}

int launch_halo_faces(const jxf_solver* h, double* prims, double* cons, int face_mask, cudaStream_t st) {
  HaloArgs a;
  a.prims = prims;
  a.cons = cons;
  a.gamma = h->cfg.gamma;
  a.fd = h->face_data;
  a.has_fd = h->has_face_data;
  long long maxcells = 0;
  for (int f = 0; f < 6; ++f) {
    a.bc[f] = ((face_mask >> f) & 1) ? h->cfg.bc[f] : JXF_BC_INACTIVE;
    for (int k = 0; k < 3; ++k) a.wall[f][k] = h->cfg.wall_velocity[f][k];
    for (int k = 0; k < 5; ++k) a.dirichlet[f][k] = h->cfg.dirichlet[f][k];
    const int ax = f >> 1;
    if (h->g.n[ax] <= 1) a.bc[f] = JXF_BC_INACTIVE;
    const int t1 = (ax == 0) ? 1 : 0, t2 = (ax == 2) ? 1 : 2;
    if (a.bc[f] != JXF_BC_INACTIVE && a.bc[f] != JXF_BC_NEIGHBOR)
      maxcells = std::max(maxcells, (long long)h->g.nh * h->g.n[t1] * h->g.n[t2]);
  }
  if (maxcells > 0) {
    const int bx = (int)std::min<long long>((maxcells + 127) / 128, 148 * 16);
    ProfScope prof(h, JXF_PROFILE_HALO, st);
    halo_fill_kernel<<<dim3(bx, 6), 128, 0, st>>>(h->g, a);
    return check_launch("halo_fill");
  }
  return JXF_OK;
}

extern "C" int jxf_halo_fill(jxf_handle h, double* prims, double* cons, void* stream) {
  if (!h || !prims || !cons) return fail(JXF_ERR_BAD_ARG, "jxf_halo_fill: null argument");
  if (int rc = launch_halo_faces(h, prims, cons, 0x3f, (cudaStream_t)stream)) return rc;
  if (dissipative(h)) return jxf_halo_fill_edges(h, prims, cons, stream);
  return JXF_OK;
}

extern "C" int jxf_stage(jxf_handle h, int stage, const double* prims_in, double* prims_out, const double* cons_in,
                         const double* cons_n, double* cons_out, double* rhs_scratch, const double* dt_dev,
                         double* red_dev, int reduce, int fill_halo, void* stream) {
  return jxf_stage_tail(h, stage, 0, prims_in, prims_out, cons_in, cons_n, cons_out, rhs_scratch, dt_dev, red_dev, reduce,
                        fill_halo, stream);
}

extern "C" int jxf_stage_tail(jxf_handle h, int stage, int first_axis_index, const double* prims_in, double* prims_out,
                              const double* cons_in, const double* cons_n, double* cons_out, double* rhs_scratch,
                              const double* dt_dev, double* red_dev, int reduce, int fill_halo, void* stream) {
  if (!h || !prims_in || !prims_out || !cons_in || !cons_out || !dt_dev)
    return fail(JXF_ERR_BAD_ARG, "jxf_stage: null argument");
  if (first_axis_index < 0 || first_axis_index >= h->n_active)
    return fail(JXF_ERR_BAD_ARG, "jxf_stage_tail: first_axis_index %d out of range", first_axis_index);
  if (stage < 0 || stage >= h->stages) return fail(JXF_ERR_BAD_ARG, "jxf_stage: stage %d out of range", stage);
  if (stage > 0 && !cons_n) return fail(JXF_ERR_BAD_ARG, "jxf_stage: cons_n required for stage > 0");
  if (prims_in == prims_out) return fail(JXF_ERR_BAD_ARG, "jxf_stage: prims_out must not alias prims_in");
  const bool diss = dissipative(h);
  if ((h->n_active > 1 || diss) && !rhs_scratch) return fail(JXF_ERR_BAD_ARG, "jxf_stage: rhs_scratch required");
  if (reduce && !red_dev) return fail(JXF_ERR_BAD_ARG, "jxf_stage: red_dev required when reduce != 0");
  if (diss) {
    // viscous + heat flux divergence of all axes first: rhs = D_x + D_y + D_z, the convective sweeps add to it
    if (first_axis_index != 0) return fail(JXF_ERR_UNSUPPORTED, "jxf_stage_tail: partial stages are not available with the viscous / heat flux");
    for (int k = 0; k < h->n_active; ++k) {
      int rc = dissipative_sweep(h, h->active[k], prims_in, rhs_scratch, k > 0, (cudaStream_t)stream);
      if (rc) return rc;
    }
  }
  if (h->cfg.no_convective_flux) {
    // no sweep to carry the fused epilogue: plain update kernel, then the halo kernels
    UpdateArgs u;
    u.cons_in = cons_in; u.cons_n = cons_n; u.rhs = rhs_scratch; u.cons_out = cons_out; u.prims_out = prims_out;
    u.dt = dt_dev; u.red = red_dev;
    u.ca = h->blend[stage][0]; u.cb = h->blend[stage][1]; u.dt_mult = h->dt_mult[stage]; u.gamma = h->cfg.gamma;
    for (int q = 0; q < 3; ++q) u.gravity[q] = h->cfg.gravity[q];
    u.blend = stage > 0; u.reduce = reduce ? 1 : 0; u.active_mask = h->active_mask; u.volume_force = h->cfg.volume_force;
    const int bx = (int)std::min<long long>((h->g.rvst + 255) / 256, 148 * 8);
    {
      ProfScope prof(h, JXF_PROFILE_OTHER, (cudaStream_t)stream);
      update_stage_kernel<<<bx, 256, 0, (cudaStream_t)stream>>>(h->g, u);
    }
    int rc = check_launch("update_stage");
    if (rc) return rc;
    return fill_halo ? jxf_halo_fill(h, prims_out, cons_out, stream) : JXF_OK;
  }
  for (int k = first_axis_index; k < h->n_active; ++k) {
    const int axis = h->order[k];
    const bool last = (k == h->n_active - 1);
    SweepArgs a = base_args(h, axis, prims_in, rhs_scratch);
    a.fl.dt = dt_dev;
    int rc;
    if (!last) {
      a.accumulate = k > 0 || diss;
      rc = dispatch_axis(h, axis, a, 0, (cudaStream_t)stream);
    } else {
      a.cons_in = cons_in;
      a.cons_n = cons_n;
      a.cons_out = cons_out;
      a.prims_out = prims_out;
      a.dt = dt_dev;
      a.red = red_dev;
      a.blend = stage > 0;
      a.ca = h->blend[stage][0];
      a.cb = h->blend[stage][1];
      a.dt_mult = h->dt_mult[stage];
      a.has_prev = k > 0 || diss;
      a.reduce = reduce ? 1 : 0;
      a.fuse_halo = fill_halo ? 1 : 0;      // outer-BC halo images written by the epilogue itself
      a.nh = h->cfg.nh;
      for (int f = 0; f < 6; ++f) {
        a.bc[f] = (h->g.n[f >> 1] > 1) ? h->cfg.bc[f] : JXF_BC_INACTIVE;
        for (int q = 0; q < 3; ++q) a.wall[f][q] = h->cfg.wall_velocity[f][q];
        for (int q = 0; q < 5; ++q) a.dirichlet[f][q] = h->cfg.dirichlet[f][q];
      }
      a.volume_force = h->cfg.volume_force;
      for (int q = 0; q < 3; ++q) a.gravity[q] = h->cfg.gravity[q];
      a.face_data = h->face_data;
      a.has_face_data = h->has_face_data;
      for (int q = 0; q < 3; ++q) a.n_phys[q] = h->g.n[q];
      for (int f = 0; f < 6; ++f) {      // direct halo stores into the neighbours' buffers (jxf_set_peer_halo)
        const bool on = fill_halo && a.bc[f] == JXF_BC_NEIGHBOR && h->peer_prims[f] && h->peer_cons[f];
        a.peer_prims[f] = on ? h->peer_prims[f] : nullptr;
        a.peer_cons[f] = on ? h->peer_cons[f] : nullptr;
      }
      rc = dispatch_axis(h, axis, a, 1, (cudaStream_t)stream);
    }
    if (rc) return rc;
  }
  // the next stage's dissipative stencils read edge halos (halo_manager.py:119-129)
  if (diss && fill_halo) return jxf_halo_fill_edges(h, prims_out, cons_out, stream);
  return JXF_OK;
}

// ---------------------------------------------------------------------------
// In-place stage on THREE full-size buffers (prims, U, U^n) + two slab-sized rhs accumulators: the memory plan for
// blocks whose five-buffer plan does not fit the device (1024^3 on one B200: 3 x 44.2 GB + 2 x 2.7 GB).
//
// The block is cut into slabs of `slab_planes` x planes.  The only reader of a slab's OLD primitives from outside
// the slab is the x sweep of its two neighbours, so the x sweep runs ONE SLAB AHEAD of the y sweep / z sweep +
// epilogue, which then update the slab's primitives and conservatives in place:
//     X(0), X(1), [Y(0), Z+epi(0)], X(2), [Y(1), Z+epi(1)], ...        (stream order)
// Inside the z sweep the rows kernel waits for the next staged window before it stores (sweep_rows, a.inplace).
// Fused halo images are written into the same buffers; the two whose OLD halo values are still needed later in the
// stage -- PERIODIC east (read by the x sweep of the last slab) and PERIODIC top (read by the last window of the
// same row) -- are deferred to one halo_fill launch after the last slab.  Same arithmetic as jxf_stage.
// Reference semantics kept: RK3.py:27-62, space_solver.py:266-314 (rhs = ((0 + x) + y) + z per cell).
// ---------------------------------------------------------------------------
extern "C" int64_t jxf_rhs_slab_elems(jxf_handle h, int slab_planes) {
  if (!h || slab_planes < 1) return -1;
  return 5LL * std::min(slab_planes, h->g.n[0]) * h->g.n[1] * h->g.n[2];
}

extern "C" int jxf_stage_inplace(jxf_handle h, int stage, double* prims, const double* cons_in, const double* cons_n,
                                 double* cons_out, double* rhs_slabs, int slab_planes, const double* dt_dev,
                                 double* red_dev, int reduce, int fill_halo, void* stream) {
  if (!h || !prims || !cons_in || !cons_out || !rhs_slabs || !dt_dev) return fail(JXF_ERR_BAD_ARG, "jxf_stage_inplace: null argument");
  if (stage < 0 || stage >= h->stages) return fail(JXF_ERR_BAD_ARG, "jxf_stage_inplace: stage %d out of range", stage);
  if (stage > 0 && !cons_n) return fail(JXF_ERR_BAD_ARG, "jxf_stage_inplace: cons_n required for stage > 0");
  if (reduce && !red_dev) return fail(JXF_ERR_BAD_ARG, "jxf_stage_inplace: red_dev required when reduce != 0");
  if (h->n_active != 3 || h->order[2] != 2)
    return fail(JXF_ERR_UNSUPPORTED, "jxf_stage_inplace: 3-D blocks with the reference sweep order only");
  if (dissipative(h) || h->cfg.no_convective_flux)
    return fail(JXF_ERR_UNSUPPORTED, "jxf_stage_inplace: convective flux only");
  const Geom& g = h->g;
  const int nx = g.n[0];
  const int P = std::min(slab_planes, nx);
  if (P < 2 * 3 && P < nx) return fail(JXF_ERR_BAD_ARG, "jxf_stage_inplace: slab_planes=%d too thin", slab_planes);
  const int nslabs = (nx + P - 1) / P;
  const long long slab_rvst = (long long)P * g.n[1] * g.n[2];
  double* slab_buf[2] = {rhs_slabs, rhs_slabs + 5 * slab_rvst};
  cudaStream_t st = (cudaStream_t)stream;

  auto sweep_x = [&](int sidx) -> int {
    const int x0 = sidx * P, x1 = std::min(nx, x0 + P);
    // the kernel indexes the rhs with the GLOBAL x index: shift the slab accumulator's base accordingly
    SweepArgs a = base_args(h, 0, prims, slab_buf[sidx & 1] - (long long)x0 * g.rst[0]);
    a.fl.dt = dt_dev;
    a.accumulate = 0;
    a.range_lo = x0;
    a.range_hi = x1;
    a.rvst_slab = slab_rvst;
    return dispatch_axis(h, 0, a, 0, st);
  };
  int rc = sweep_x(0);
  if (rc) return rc;
  bool defer[6] = {false, false, false, false, false, false};
  for (int sidx = 0; sidx < nslabs; ++sidx) {
    if (sidx + 1 < nslabs && (rc = sweep_x(sidx + 1))) return rc;
    const int x0 = sidx * P, x1 = std::min(nx, x0 + P);
    {   // y sweep of the slab
      SweepArgs a = base_args(h, 1, prims, slab_buf[sidx & 1]);
      a.fl.dt = dt_dev;
      a.accumulate = 1;
      a.sub_lo = x0; a.sub_n = x1 - x0; a.rvst_slab = slab_rvst;
      if ((rc = dispatch_axis(h, 1, a, 0, st))) return rc;
    }
    {   // z sweep + epilogue of the slab, in place
      SweepArgs a = base_args(h, 2, prims, slab_buf[sidx & 1]);
      a.fl.dt = dt_dev;
      a.sub_lo = x0; a.sub_n = x1 - x0; a.rvst_slab = slab_rvst;
      a.cons_in = cons_in; a.cons_n = cons_n; a.cons_out = cons_out; a.prims_out = prims;
      a.dt = dt_dev; a.red = red_dev;
      a.blend = stage > 0; a.ca = h->blend[stage][0]; a.cb = h->blend[stage][1]; a.dt_mult = h->dt_mult[stage];
      a.has_prev = 1; a.reduce = reduce ? 1 : 0; a.fuse_halo = fill_halo ? 1 : 0; a.nh = h->cfg.nh; a.inplace = 1;
      for (int f = 0; f < 6; ++f) {
        a.bc[f] = h->cfg.bc[f];
        for (int q = 0; q < 3; ++q) a.wall[f][q] = h->cfg.wall_velocity[f][q];
        for (int q = 0; q < 5; ++q) a.dirichlet[f][q] = h->cfg.dirichlet[f][q];
      }
      // east (face 0) / top (face 4) PERIODIC images overwrite halos the stage still reads: deferred
      for (int f : {0, 4})
        if (a.bc[f] == JXF_BC_PERIODIC) { a.bc[f] = JXF_BC_NEIGHBOR; defer[f] = true; }
      a.volume_force = h->cfg.volume_force;
      for (int q = 0; q < 3; ++q) a.gravity[q] = h->cfg.gravity[q];
      a.face_data = h->face_data;
      a.has_face_data = h->has_face_data;
      for (int q = 0; q < 3; ++q) a.n_phys[q] = h->g.n[q];
      if ((rc = dispatch_axis(h, 2, a, 1, st))) return rc;
    }
  }
  if (fill_halo && (defer[0] || defer[4])) {
    HaloArgs a;
    memset(&a, 0, sizeof(a));
    a.prims = prims; a.cons = cons_out; a.gamma = h->cfg.gamma;
    long long maxcells = 0;
    for (int f = 0; f < 6; ++f) {
      a.bc[f] = defer[f] ? JXF_BC_PERIODIC : JXF_BC_INACTIVE;
      const int ax = f >> 1, t1 = (ax == 0) ? 1 : 0, t2 = (ax == 2) ? 1 : 2;
      if (defer[f]) maxcells = std::max(maxcells, (long long)g.nh * g.n[t1] * g.n[t2]);
    }
    const int bx = (int)std::min<long long>((maxcells + 127) / 128, 148 * 16);
    ProfScope prof(h, JXF_PROFILE_HALO, st);
    halo_fill_kernel<<<dim3(bx, 6), 128, 0, st>>>(h->g, a);
    if ((rc = check_launch("halo_fill (deferred)"))) return rc;
  }
  return JXF_OK;
}

extern "C" int jxf_step_fused(jxf_handle h, double* prims_a, double* prims_b, double* cons_a, double* cons_b,
                              double* rhs_scratch, double* dt_dev, double* time_dev, double* red_dev,
                              double* info_dev, int fill_halo, void* stream) {
  if (!h || !prims_a || !prims_b || !cons_a || !cons_b || !dt_dev || !red_dev)
    return fail(JXF_ERR_BAD_ARG, "jxf_step_fused: null argument");
  double* pr[2] = {prims_a, prims_b};
  int cur = 0;
  for (int k = 0; k < h->stages; ++k) {
    const bool last = (k == h->stages - 1);
    const double* cin = (k == 0) ? cons_a : cons_b;
    double* cout = last ? cons_a : cons_b;
    int rc = jxf_stage(h, k, pr[cur], pr[cur ^ 1], cin, cons_a, cout, rhs_scratch, dt_dev, red_dev, last ? 1 : 0,
                       fill_halo, stream);
    if (rc) return rc;
    cur ^= 1;
  }
  int rc = jxf_finish_step(h, red_dev, dt_dev, time_dev, info_dev, stream);
  if (rc) return rc;
  return cur;
}

extern "C" int jxf_prims_from_cons(jxf_handle h, const double* cons, double* prims, void* stream) {
  if (!h || !cons || !prims) return fail(JXF_ERR_BAD_ARG, "jxf_prims_from_cons: null argument");
  const int bx = (int)std::min<long long>((h->g.vst + 255) / 256, 148 * 8);
  ProfScope prof(h, JXF_PROFILE_OTHER, (cudaStream_t)stream);
  prims_from_cons_kernel<<<bx, 256, 0, (cudaStream_t)stream>>>(cons, prims, h->g.vst, h->cfg.gamma);
  return check_launch("prims_from_cons");
}

extern "C" int jxf_cons_from_prims(jxf_handle h, const double* prims, double* cons, void* stream) {
  if (!h || !cons || !prims) return fail(JXF_ERR_BAD_ARG, "jxf_cons_from_prims: null argument");
  const int bx = (int)std::min<long long>((h->g.vst + 255) / 256, 148 * 8);
  ProfScope prof(h, JXF_PROFILE_OTHER, (cudaStream_t)stream);
  cons_from_prims_kernel<<<bx, 256, 0, (cudaStream_t)stream>>>(prims, cons, h->g.vst, h->cfg.gamma);
  return check_launch("cons_from_prims");
}

extern "C" int jxf_reduce(jxf_handle h, const double* prims, double* red_dev, void* stream) {
  if (!h || !prims || !red_dev) return fail(JXF_ERR_BAD_ARG, "jxf_reduce: null argument");
  const long long total = h->g.rvst;
  const int bx = (int)std::min<long long>((total + 255) / 256, 148 * 8);
  ProfScope prof(h, JXF_PROFILE_OTHER, (cudaStream_t)stream);
  reduce_kernel<<<bx, 256, 0, (cudaStream_t)stream>>>(h->g, prims, red_dev, h->cfg.gamma, h->active_mask);
  return check_launch("reduce");
}

extern "C" int jxf_reduce_reset(jxf_handle h, double* red_dev, void* stream) {
  if (!h || !red_dev) return fail(JXF_ERR_BAD_ARG, "jxf_reduce_reset: null argument");
  ProfScope prof(h, JXF_PROFILE_OTHER, (cudaStream_t)stream);
  reduce_reset_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(red_dev);
  return check_launch("reduce_reset");
}

extern "C" int jxf_finish_step(jxf_handle h, double* red_dev, double* dt_dev, double* time_dev, double* info_dev,
                               void* stream) {
  if (!h || !red_dev || !dt_dev) return fail(JXF_ERR_BAD_ARG, "jxf_finish_step: null argument");
  ProfScope prof(h, JXF_PROFILE_OTHER, (cudaStream_t)stream);
  DtLimits lim;
  lim.visc = h->cfg.viscous_flux;
  lim.heat = h->cfg.heat_flux;
  lim.mu = h->cfg.dynamic_viscosity;
  lim.lambda = h->cfg.thermal_conductivity;
  lim.cp = h->cfg.gamma / (h->cfg.gamma - 1.0) * h->cfg.gas_constant;      // ideal_gas.py:33
  finish_step_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(red_dev, dt_dev, time_dev, info_dev, h->cfg.dx_min, h->cfg.cfl,
                                                        h->cfg.fixed_dt, lim);
  return check_launch("finish_step");
}

extern "C" int jxf_integrate_stage(jxf_handle h, int stage, const double* cons, const double* cons_n, const double* rhs,
                                   double dt, double* cons_out, void* stream) {
  if (!h || !cons || !rhs || !cons_out) return fail(JXF_ERR_BAD_ARG, "jxf_integrate_stage: null argument");
  if (stage < 0 || stage >= h->stages) return fail(JXF_ERR_BAD_ARG, "jxf_integrate_stage: stage %d out of range", stage);
  if (stage > 0 && !cons_n) return fail(JXF_ERR_BAD_ARG, "jxf_integrate_stage: cons_n required for stage > 0");
  const int bx = (int)std::min<long long>((h->g.vst + 255) / 256, 148 * 16);
  ProfScope prof(h, JXF_PROFILE_OTHER, (cudaStream_t)stream);
  integrate_stage_kernel<<<bx, 256, 0, (cudaStream_t)stream>>>(h->g, cons, cons_n, rhs, cons_out, h->blend[stage][0],
                                                               h->blend[stage][1], stage > 0, dt * h->dt_mult[stage]);
  return check_launch("integrate_stage");
}

// transverse range of a face slab: interior, widened by the nh halo cells on the sides named in ext_mask
// (bit 0: low side of the slower transverse axis, bit 1: its high side, bit 2 / 3: the faster transverse axis)
static void slab_ranges(const jxf_solver* h, int face, int ext_mask, int& lo1, int& n1, int& lo2, int& n2) {
  const int ax = face >> 1;
  const int t1 = (ax == 0) ? 1 : 0, t2 = (ax == 2) ? 1 : 2;
  const Geom& g = h->g;
  lo1 = g.off[t1]; n1 = g.n[t1];
  lo2 = g.off[t2]; n2 = g.n[t2];
  if (g.n[t1] > 1) {
    if (ext_mask & 1) { lo1 -= g.nh; n1 += g.nh; }
    if (ext_mask & 2) n1 += g.nh;
  }
  if (g.n[t2] > 1) {
    if (ext_mask & 4) { lo2 -= g.nh; n2 += g.nh; }
    if (ext_mask & 8) n2 += g.nh;
  }
}

extern "C" int64_t jxf_face_slab_elems_n(jxf_handle h, int face, int ext_mask, int layers) {
  if (!h || face < 0 || face > 5 || layers < 1 || layers > h->g.nh) return -1;
  int lo1, n1, lo2, n2;
  slab_ranges(h, face, ext_mask, lo1, n1, lo2, n2);
  return 5LL * layers * n1 * n2;
}

extern "C" int64_t jxf_face_slab_elems_ext(jxf_handle h, int face, int ext_mask) {
  return h ? jxf_face_slab_elems_n(h, face, ext_mask, h->g.nh) : -1;
}

extern "C" int64_t jxf_face_slab_elems(jxf_handle h, int face) { return jxf_face_slab_elems_ext(h, face, 0); }

static int face_slab(jxf_handle h, int face, int ext_mask, double* prims, double* cons, double* slab, int unpack, void* stream,
                     int layers = 0) {
  if (!h || !prims || !slab || (unpack && !cons)) return fail(JXF_ERR_BAD_ARG, "jxf_(un)pack_face: null argument");
  if (layers == 0) layers = h->g.nh;
  if (layers < 1 || layers > h->g.nh) return fail(JXF_ERR_BAD_ARG, "jxf_(un)pack_face: layers=%d outside [1, halo_cells]", layers);
  if (face < 0 || face > 5 || h->g.n[face >> 1] <= 1) return fail(JXF_ERR_BAD_ARG, "jxf_(un)pack_face: face %d not active", face);
  if (ext_mask < 0 || ext_mask > 15) return fail(JXF_ERR_BAD_ARG, "jxf_(un)pack_face: ext_mask %d", ext_mask);
  FaceArgs a;
  a.prims = prims;
  a.cons = cons;
  a.slab = slab;
  a.gamma = h->cfg.gamma;
  a.face = face;
  a.unpack = unpack;
  a.layers = layers;
  slab_ranges(h, face, ext_mask, a.lo1, a.n1, a.lo2, a.n2);
  const long long total = (long long)layers * a.n1 * a.n2;
  const int bx = (int)std::min<long long>((total + 127) / 128, 148 * 16);
  ProfScope prof(h, JXF_PROFILE_OTHER, (cudaStream_t)stream);
  face_slab_kernel<<<bx, 128, 0, (cudaStream_t)stream>>>(h->g, a);
  return check_launch("face_slab");
}

extern "C" int jxf_pack_face(jxf_handle h, int face, const double* prims, double* slab, void* stream) {
  return face_slab(h, face, 0, const_cast<double*>(prims), nullptr, slab, 0, stream);
}
extern "C" int jxf_unpack_face(jxf_handle h, int face, const double* slab, double* prims, double* cons, void* stream) {
  return face_slab(h, face, 0, prims, cons, const_cast<double*>(slab), 1, stream);
}
extern "C" int jxf_pack_face_ext(jxf_handle h, int face, int ext_mask, const double* prims, double* slab, void* stream) {
  return face_slab(h, face, ext_mask, const_cast<double*>(prims), nullptr, slab, 0, stream);
}
extern "C" int jxf_unpack_face_ext(jxf_handle h, int face, int ext_mask, const double* slab, double* prims, double* cons,
                                   void* stream) {
  return face_slab(h, face, ext_mask, prims, cons, const_cast<double*>(slab), 1, stream);
}

extern "C" int jxf_pack_face_n(jxf_handle h, int face, int ext_mask, int layers, const double* prims, double* slab, void* stream) {
  return face_slab(h, face, ext_mask, const_cast<double*>(prims), nullptr, slab, 0, stream, layers);
}
extern "C" int jxf_unpack_face_n(jxf_handle h, int face, int ext_mask, int layers, const double* slab, double* prims,
                                 double* cons, void* stream) {
  return face_slab(h, face, ext_mask, prims, cons, const_cast<double*>(slab), 1, stream, layers);
}

extern "C" int jxf_debug_face_flux(int axis, int recon, int riemann, const double* windows, int64_t n, double gamma,
                                   double* flux, void* stream) {
  if (!windows || !flux || n <= 0) return fail(JXF_ERR_BAD_ARG, "jxf_debug_face_flux: bad argument");
  const unsigned bx = (unsigned)((n + 127) / 128);
  cudaStream_t st = (cudaStream_t)stream;
  (void)bx; (void)st; (void)axis; (void)riemann; (void)gamma;
  // stencils other than the two WENO5 forms: the generic instantiations + the stencil id in the option word
  const int stencil = recon >> 1;
  if (recon < 0 || stencil > JXF_STENCIL_TENO6A) return fail(JXF_ERR_BAD_ARG, "jxf_debug_face_flux: unknown variant");
  const int opt = stencil >= JXF_STENCIL_WENO1 ? stencil_bits(stencil) | ((recon & 1) << 19) : 0;
  recon = (recon & 1) + 2 * std::min(stencil, (int)STENCIL_GENERIC);
  (void)opt;
#define JXF_DBG_CASE(A, R, S)                                                                     \
  if (axis == A && recon == R && riemann == S) {                                                  \
    face_flux_debug_kernel<A, R, S><<<bx, 128, 0, st>>>(windows, (long long)n, gamma, flux, opt);  \
    return check_launch("face_flux_debug");                                                       \
  }
#ifndef JXF_TUNE_ONLY
  JXF_DBG_CASE(0, 0, 0) JXF_DBG_CASE(0, 0, 1) JXF_DBG_CASE(0, 1, 0) JXF_DBG_CASE(0, 1, 1)
  JXF_DBG_CASE(1, 0, 0) JXF_DBG_CASE(1, 0, 1) JXF_DBG_CASE(1, 1, 0) JXF_DBG_CASE(1, 1, 1)
  JXF_DBG_CASE(2, 0, 0) JXF_DBG_CASE(2, 0, 1) JXF_DBG_CASE(2, 1, 0) JXF_DBG_CASE(2, 1, 1)
  JXF_DBG_CASE(0, 2, 0) JXF_DBG_CASE(0, 3, 0) JXF_DBG_CASE(1, 2, 0) JXF_DBG_CASE(1, 3, 0)
  JXF_DBG_CASE(2, 2, 0) JXF_DBG_CASE(2, 3, 0)
  JXF_DBG_CASE(0, 4, 0) JXF_DBG_CASE(0, 5, 0) JXF_DBG_CASE(1, 4, 0) JXF_DBG_CASE(1, 5, 0)
  JXF_DBG_CASE(2, 4, 0) JXF_DBG_CASE(2, 5, 0)
#endif
#undef JXF_DBG_CASE
  return fail(JXF_ERR_BAD_ARG, "jxf_debug_face_flux: unknown variant");
}

extern "C" int jxf_debug_dispatch(jxf_handle h, int axis, int* recon_template, int* riemann_template, int* option_word) {
  if (!h || axis < 0 || axis > 2 || !recon_template || !riemann_template || !option_word)
    return fail(JXF_ERR_BAD_ARG, "jxf_debug_dispatch: bad argument");
  *recon_template = recon_template_of(h);
  *riemann_template = riemann_template_of(h);
  *option_word = base_args(h, axis, nullptr, nullptr).limiter;
  return JXF_OK;
}

extern "C" int jxf_debug_math(const double* x, int64_t n, double* out, void* stream) {
  if (!x || !out || n <= 0) return fail(JXF_ERR_BAD_ARG, "jxf_debug_math: bad argument");
#ifndef JXF_REFERENCE_ORDER
  math_debug_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(x, (long long)n, out);
  return check_launch("math_debug");
#else
  return fail(JXF_ERR_UNSUPPORTED, "jxf_debug_math: built with JXF_REFERENCE_ORDER");
#endif
}

extern "C" int jxf_fp64_probe(double* scratch, int iters, int64_t* n_fma, void* stream) {
  if (!scratch || iters <= 0) return fail(JXF_ERR_BAD_ARG, "jxf_fp64_probe: bad argument");
  const int blocks = 148 * 8, threads = 256;
  fp64_probe_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(scratch, iters);
  if (n_fma) *n_fma = (int64_t)blocks * threads * 8LL * iters;
  return check_launch("fp64_probe");
}
