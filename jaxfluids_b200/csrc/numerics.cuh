// numerics.cuh -- per-face / per-cell fp64 device functions of the convective path.
//
// Everything here is a pure function of registers; the sweep kernels in
// jxf_b200.cu decide how the 6-cell windows get there.  Formulas follow the
// reference (citations: file:line under /root/reference/src/jaxfluids/), the
// evaluation is re-associated / reciprocal-hoisted where that stays far inside
// the 1e-12 parity tolerance (see DESIGN.md "arithmetic").
#pragma once
#include <cuda_runtime.h>

namespace jxf {

constexpr double kStencilEps = 1e-30;                 // config/precision.py:53
constexpr double kEps = 2.220446049250313e-16;        // config/precision.py:44-55

enum { RECON_PRIMITIVE = 0, RECON_CHAR_PRIMITIVE = 1 };
enum { RIEMANN_HLLC = 0, RIEMANN_RUSANOV = 1 };

// velocity_minor_axes, equation_information.py:110
template <int A> struct AxisIds;
template <> struct AxisIds<0> { static constexpr int un = 1, t0 = 2, t1 = 3; };
template <> struct AxisIds<1> { static constexpr int un = 2, t0 = 3, t1 = 1; };
template <> struct AxisIds<2> { static constexpr int un = 3, t0 = 1, t1 = 2; };

// ---------------------------------------------------------------------------
// WENO5-Z  (stencils/reconstruction/shock_capturing/weno5_base.py:34-51,
//           weno/weno5_z.py:32-52).  (a,b,c,d,e) = cells i-2..i+2 for the left
//           state, mirrored (i+3..i-1) for the right state.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double weno5z(double a, double b, double c, double d, double e) {
  const double s0 = a - 2.0 * b + c;
  const double q0 = a - 4.0 * b + 3.0 * c;
  const double s1 = b - 2.0 * c + d;
  const double q1 = b - d;
  const double s2 = c - 2.0 * d + e;
  const double q2 = 3.0 * c - 4.0 * d + e;
  const double beta0 = (13.0 / 12.0) * (s0 * s0) + 0.25 * (q0 * q0);
  const double beta1 = (13.0 / 12.0) * (s1 * s1) + 0.25 * (q1 * q1);
  const double beta2 = (13.0 / 12.0) * (s2 * s2) + 0.25 * (q2 * q2);
  const double tau5 = fabs(beta0 - beta2);
  const double alpha0 = 0.1 * (1.0 + tau5 / (beta0 + kStencilEps));
  const double alpha1 = 0.6 * (1.0 + tau5 / (beta1 + kStencilEps));
  const double alpha2 = 0.3 * (1.0 + tau5 / (beta2 + kStencilEps));
  const double inv = 1.0 / (alpha0 + alpha1 + alpha2);
  const double p0 = (1.0 / 3.0) * a + (-7.0 / 6.0) * b + (11.0 / 6.0) * c;
  const double p1 = (-1.0 / 6.0) * b + (5.0 / 6.0) * c + (1.0 / 3.0) * d;
  const double p2 = (1.0 / 3.0) * c + (5.0 / 6.0) * d + (-1.0 / 6.0) * e;
  return (alpha0 * inv) * p0 + (alpha1 * inv) * p1 + (alpha2 * inv) * p2;
}

__device__ __forceinline__ void weno5z_lr(const double (&q)[6], double& left, double& right) {
  left = weno5z(q[0], q[1], q[2], q[3], q[4]);
  right = weno5z(q[5], q[4], q[3], q[2], q[1]);
}

// ---------------------------------------------------------------------------
// EOS / variable transforms (materials/single_materials/ideal_gas.py:69-88,
// equation_manager.py:93-101, 164-171, 237-252)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void cons_from_prims(const double (&p)[5], double gamma, double (&c)[5]) {
  const double e = p[4] / (p[0] * (gamma - 1.0));
  c[0] = p[0];
  c[1] = p[0] * p[1];
  c[2] = p[0] * p[2];
  c[3] = p[0] * p[3];
  c[4] = p[0] * (0.5 * ((p[1] * p[1] + p[2] * p[2]) + p[3] * p[3]) + e);
}

__device__ __forceinline__ void prims_from_cons(const double (&c)[5], double gamma, double (&p)[5]) {
  const double one_rho = 1.0 / c[0];
  p[0] = c[0];
  p[1] = c[1] * one_rho;
  p[2] = c[2] * one_rho;
  p[3] = c[3] * one_rho;
  const double e = c[4] * one_rho - 0.5 * ((p[1] * p[1] + p[2] * p[2]) + p[3] * p[3]);
  p[4] = (gamma - 1.0) * e * c[0];
}

template <int A>
__device__ __forceinline__ void physical_flux(const double (&p)[5], const double (&c)[5], double (&f)[5]) {
  const double m = c[1 + A];
  f[0] = m;
  f[1] = m * p[1];
  f[2] = m * p[2];
  f[3] = m * p[3];
  f[1 + A] += p[4];
  f[4] = p[1 + A] * (c[4] + p[4]);
}

// ---------------------------------------------------------------------------
// Reconstruction of the left/right face states from the 6-cell window
// w[var][k], k = cells i-2..i+3 around the face between i and i+1.
// PRIMITIVE:      high_order_godunov.py:267-280
// CHAR-PRIMITIVE: high_order_godunov.py:298-316 with eigendecomposition.py
//                 :139-148,215-231 (frozen arithmetic state), :425-431, :517-521.
// ---------------------------------------------------------------------------
template <int A, int RECON>
__device__ __forceinline__ void reconstruct(const double (&w)[5][6], double gamma,
                                            double (&pl)[5], double (&pr)[5]) {
  using Id = AxisIds<A>;
  if (RECON == RECON_PRIMITIVE) {
#pragma unroll
    for (int v = 0; v < 5; ++v) weno5z_lr(w[v], pl[v], pr[v]);
  } else {
    const double rho_ave = 0.5 * (w[0][2] + w[0][3]);
    const double p_ave = 0.5 * (w[4][2] + w[4][3]);
    const double c_ave = sqrt(gamma * p_ave / rho_ave);
    const double cc_ave = c_ave * c_ave;
    const double k_u = 0.5 / c_ave;
    const double k_p = 0.5 / (cc_ave * rho_ave);
    const double k_cc = 1.0 / cc_ave;
    double q[6], l0, r0, l1, r1, l4, r4;
#pragma unroll
    for (int k = 0; k < 6; ++k) q[k] = -k_u * w[Id::un][k] + k_p * w[4][k];
    weno5z_lr(q, l0, r0);
#pragma unroll
    for (int k = 0; k < 6; ++k) q[k] = w[0][k] - k_cc * w[4][k];
    weno5z_lr(q, l1, r1);
#pragma unroll
    for (int k = 0; k < 6; ++k) q[k] = k_u * w[Id::un][k] + k_p * w[4][k];
    weno5z_lr(q, l4, r4);
    weno5z_lr(w[Id::t0], pl[Id::t0], pr[Id::t0]);
    weno5z_lr(w[Id::t1], pl[Id::t1], pr[Id::t1]);
    const double ccr = cc_ave * rho_ave;
    pl[0] = rho_ave * (l0 + l4) + l1;
    pl[Id::un] = c_ave * (-l0 + l4);
    pl[4] = ccr * (l0 + l4);
    pr[0] = rho_ave * (r0 + r4) + r1;
    pr[Id::un] = c_ave * (-r0 + r4);
    pr[4] = ccr * (r0 + r4);
  }
}

// ---------------------------------------------------------------------------
// Riemann solvers
// ---------------------------------------------------------------------------
// HLLC with Einfeldt signal speeds: solvers/riemann_solvers/HLLC.py:41-126,
// signal_speeds.py:109-133 (Einfeldt), :159-199 (S*).
template <int A, bool LEFT>
__device__ __forceinline__ void hllc_star_flux(const double (&p)[5], const double (&c)[5],
                                               double S_K, double S_star, double (&fs)[5]) {
  using Id = AxisIds<A>;
  const double dK = S_K - p[Id::un];
  const double pre = dK / (S_K - S_star) * p[0];
  double us[5];
  us[0] = pre;
  us[Id::un] = pre * S_star;
  us[Id::t0] = pre * p[Id::t0];
  us[Id::t1] = pre * p[Id::t1];
  us[4] = pre * (c[4] / c[0] + (S_star - p[Id::un]) * (S_star + p[4] / p[0] / dK));
  double f[5];
  physical_flux<A>(p, c, f);
  const double S = LEFT ? fmin(S_K, 0.0) : fmax(S_K, 0.0);
#pragma unroll
  for (int v = 0; v < 5; ++v) fs[v] = f[v] + S * (us[v] - c[v]);
}

template <int A, int RIEMANN>
__device__ __forceinline__ void riemann_flux(const double (&pl)[5], const double (&pr)[5],
                                             double gamma, double (&F)[5]) {
  using Id = AxisIds<A>;
  double cl[5], cr[5];
  cons_from_prims(pl, gamma, cl);
  cons_from_prims(pr, gamma, cr);
  const double aL = sqrt(gamma * pl[4] / pl[0]);
  const double aR = sqrt(gamma * pr[4] / pr[0]);
  const double uL = pl[Id::un], uR = pr[Id::un];
  if (RIEMANN == RIEMANN_HLLC) {
    const double sL = sqrt(pl[0]), sR = sqrt(pr[0]);
    const double one_dens = 1.0 / (sL + sR);
    const double eta2 = 0.5 * sL * sR * one_dens * one_dens;
    const double u_bar = (sL * uL + sR * uR) * one_dens;
    const double du = uR - uL;
    const double d_bar = sqrt((sL * aL * aL + sR * aR * aR) * one_dens + eta2 * (du * du));
    const double S_L = fmin(u_bar - d_bar, uL - aL);
    const double S_R = fmax(u_bar + d_bar, uR + aR);
    const double dL = pl[0] * (S_L - uL);
    const double dR = pr[0] * (S_R - uR);
    const double S_star = ((pr[4] - pl[4]) + (uL * dL - uR * dR)) / (dL - dR);
    double fL[5], fR[5];
    hllc_star_flux<A, true>(pl, cl, S_L, S_star, fL);
    hllc_star_flux<A, false>(pr, cr, S_R, S_star, fR);
    const double sgn = (S_star > 0.0) ? 1.0 : ((S_star < 0.0) ? -1.0 : 0.0);   // jnp.sign
    const double wl = 0.5 * (1.0 + sgn), wr = 0.5 * (1.0 - sgn);
#pragma unroll
    for (int v = 0; v < 5; ++v) F[v] = wl * fL[v] + wr * fR[v];
  } else {
    // Rusanov: solvers/riemann_solvers/Rusanov.py:25-47
    const double alpha = fmax(fabs(uL) + aL, fabs(uR) + aR);
    double fl[5], fr[5];
    physical_flux<A>(pl, cl, fl);
    physical_flux<A>(pr, cr, fr);
#pragma unroll
    for (int v = 0; v < 5; ++v) F[v] = 0.5 * (fl[v] + fr[v]) - 0.5 * alpha * (cr[v] - cl[v]);
  }
}

// window -> numerical flux at the face (high_order_godunov.py:117-231)
template <int A, int RECON, int RIEMANN>
__device__ __forceinline__ void face_flux(const double (&w)[5][6], double gamma, double (&F)[5]) {
  double pl[5], pr[5];
  reconstruct<A, RECON>(w, gamma, pl, pr);
  riemann_flux<A, RIEMANN>(pl, pr, gamma, F);
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------
// Per-block reductions: max sum(|u_i|+c) (time_step_size.py:103-109), min rho,
// min p (positivity_handler.py:246-247)
// ---------------------------------------------------------------------------
struct Red {
  double max_s, min_rho, min_p;
  __device__ __forceinline__ void init() {
    max_s = 0.0;
    min_rho = __longlong_as_double(0x7ff0000000000000LL);
    min_p = min_rho;
  }
  __device__ __forceinline__ void add_cell(const double (&p)[5], double gamma, int active_mask) {
    const double c = sqrt(gamma * p[4] / p[0]);
    double s = 0.0;
    if (active_mask & 1) s += fabs(p[1]) + c;
    if (active_mask & 2) s += fabs(p[2]) + c;
    if (active_mask & 4) s += fabs(p[3]) + c;
    max_s = fmax(max_s, s);
    min_rho = fmin(min_rho, p[0]);
    min_p = fmin(min_p, p[4]);
  }
};

__device__ __forceinline__ void atomic_max_f64(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a;
  while (__longlong_as_double((long long)old) < v) {
    const unsigned long long assumed = old;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    if (old == assumed) break;
  }
}
__device__ __forceinline__ void atomic_min_f64(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a;
  while (__longlong_as_double((long long)old) > v) {
    const unsigned long long assumed = old;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    if (old == assumed) break;
  }
}

// all 32 lanes must call
__device__ __forceinline__ void red_commit(Red r, double* red_dev) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    r.max_s = fmax(r.max_s, __shfl_xor_sync(0xffffffffu, r.max_s, o));
    r.min_rho = fmin(r.min_rho, __shfl_xor_sync(0xffffffffu, r.min_rho, o));
    r.min_p = fmin(r.min_p, __shfl_xor_sync(0xffffffffu, r.min_p, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomic_max_f64(red_dev + 0, r.max_s);
    atomic_min_f64(red_dev + 1, r.min_rho);
    atomic_min_f64(red_dev + 2, r.min_p);
  }
}

#endif  // __CUDACC__

}  // namespace jxf
