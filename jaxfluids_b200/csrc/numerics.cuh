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

enum { RECON_PRIMITIVE = 0, RECON_CHAR_PRIMITIVE = 1 };   // reconstruction variable = RECON & 1
enum { STENCIL_WENO5Z = 0, STENCIL_WENO5JS = 1, STENCIL_GENERIC = 2 };   // reconstruction stencil = RECON >> 1
// The kernels' RECON template parameter carries both: RECON = variable + 2 * stencil.
// STENCIL_GENERIC: the other reconstruction stencils of the reference that fit the 6-cell window, selected at run
// time by `alt` (bits 11-14 of the face-flux option word) inside ONE extra set of kernel instantiations -- plain
// reference-order arithmetic in an out-of-line function, not a tuned path.  The ids are the C ABI's JXF_STENCIL_*.
enum { ALT_WENO5Z = 0, ALT_WENO5JS = 1,   // reference-order forms of the two tuned stencils (flux-splitting path only)
       ALT_WENO1 = 2, ALT_WENO3JS = 3, ALT_WENO3Z = 4, ALT_TENO5 = 5, ALT_WENO6CU = 6, ALT_KOREN = 7, ALT_MC = 8,
       ALT_MINMOD = 9, ALT_SUPERBEE = 10, ALT_VANALBADA = 11, ALT_VANLEER = 12, ALT_WENO3N = 13, ALT_CENTRAL2 = 14,
       ALT_TENO6 = 15, ALT_TENO5A = 16, ALT_TENO6A = 17 };   // ids >= 16: bit 22 of the option word is the fifth id bit
enum { RIEMANN_HLLC = 0, RIEMANN_RUSANOV = 1,
       // HLLC + EINFELDT with every run-time option of the face flux off (option word 0: no interpolation / flux limiter,
       // no alternative signal speed): the same arithmetic as RIEMANN_HLLC with the option branches -- and the
       // out-of-line call sites behind them -- compiled out of the sweep loops.  Tuned stencils only (RECON 0..3).
       RIEMANN_HLLC_PLAIN = 2 };
// HLLC wave-speed estimate (signal_speeds.py): a run-time option `sig` of riemann_flux (uniform branch), packed
// with the limiter mode into the `opt` argument of face_flux: opt = lim | (sig << 4) | (HLL << 8), HLL = the HLL
// Riemann solver (HLL.py) riding on the RIEMANN_RUSANOV kernel instantiations
enum { SIG_EINFELDT = 0, SIG_ARITHMETIC = 1, SIG_RUSANOV = 2, SIG_DAVIS = 3, SIG_TORO = 4 };
// Further per-face Riemann solvers of the reference, bits 15-16 of the option word (= bits 11-12 of `sig`), riding on
// the RIEMANN_RUSANOV kernel instantiations like HLL: plain reference-order arithmetic in an out-of-line function.
enum { RIEMANN_ALT_NONE = 0, RIEMANN_ALT_HLLCLM = 1 /* HLLCLM.py */, RIEMANN_ALT_AUSMP = 2 /* AUSMP.py */ };

// ---------------------------------------------------------------------------
// fast reciprocal / rsqrt / sqrt: MUFU.RCP64H / MUFU.RSQ64H seed (>= 20 bits) + Newton.
// Valid for normal, finite, positive-or-negative (rcp) / positive (rsqrt) arguments -- all
// call sites divide by densities, sound speeds, wave-speed differences and WENO weight sums.
// ---------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ double rcp_fast(double a) {
  // seed error e0 <= 2^-20 (measured 9.8e-7, tests/test_gpu_parity.py); one cubic step: x (1 + e + e^2),
  // remaining error e0^3 ~ 1e-18
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
  const double e = fma(-a, x, 1.0);
  return fma(x, fma(e, e, e), x);
}
__device__ __forceinline__ double rsqrt_fast(double a) {
  // seed error <= 2^-20; one cubic step y (1 + e/2 + 3 e^2/8), e = 1 - a y^2, remaining error ~ e^3
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  const double e = fma(-a * y, y, 1.0);
  return fma(y, e * fma(0.375, e, 0.5), y);
}
#else
__device__ __forceinline__ double rcp_fast(double a) { return 1.0 / a; }
__device__ __forceinline__ double rsqrt_fast(double a) { return 1.0 / sqrt(a); }
#endif
// sqrt(a) = a * rsqrt(a): y carries <= 2 ulp, the product <= 3 ulp -- these roots only feed wave-speed
// estimates (a_K, d_bar), whose relative error enters the flux multiplied by the (small) jump
__device__ __forceinline__ double sqrt_fast(double a, double y /* = rsqrt_fast(a) */) { return a * y; }

// Division / reciprocal / square root / 10^-n of the GENERIC (reference-order) device functions -- the other stencils,
// HLLC-LM / AUSM+ / HLL, the flux splitting, the conservative reconstruction variables.  On the device, in production
// builds: MUFU seed + Newton (<= 1.5 ulp) instead of the IEEE division / sqrt sequences (~4x fewer instructions, no slow
// path), and a table for the TENO-A cut-off 10^-n (n is a small integer by construction: ceil(.) - 1).  On the host
// (tests/hostsim) and in JXF_REFERENCE_ORDER builds: the IEEE operations, i.e. bit-identical to the reference without
// FMA contraction.  Denominators here are beta + 1e-30, sums of positive weights, densities, sound speeds, wave-speed
// differences: normal, finite, non-zero.  (The MUSCL limiters keep IEEE division: their denominators can vanish.)
#if defined(__CUDACC__) && !defined(JXF_REFERENCE_ORDER) && !defined(JXF_GENERIC_IEEE)
__device__ __forceinline__ double gdiv(double a, double b) { return a * rcp_fast(b); }
__device__ __forceinline__ double grcp(double b) { return rcp_fast(b); }
__device__ __forceinline__ double gsqrt(double a) { return a * rsqrt_fast(a); }
__device__ __forceinline__ double gpow10_neg(double n) {
  // 10^-n, n integer-valued in [0, 15] (TENO5-A: 4..9, TENO6-A: 5..10): the correctly rounded decimal literals
  const double t[16] = {1e0, 1e-1, 1e-2, 1e-3, 1e-4, 1e-5, 1e-6, 1e-7, 1e-8, 1e-9, 1e-10, 1e-11, 1e-12, 1e-13, 1e-14, 1e-15};
  const int i = (int)n;
  return (i >= 0 && i < 16 && (double)i == n) ? t[i] : pow(10.0, -n);
}
#else
__device__ __forceinline__ double gdiv(double a, double b) { return a / b; }
__device__ __forceinline__ double grcp(double b) { return 1.0 / b; }
__device__ __forceinline__ double gsqrt(double a) { return sqrt(a); }
__device__ __forceinline__ double gpow10_neg(double n) { return pow(10.0, -n); }
#endif

// signal_speeds.py:10-69, :135-157 with estimate_pressure :201-214 -- the simple estimates, reference order.
// OUT OF LINE: the tuned path is EINFELDT; keeping these (IEEE sqrt / divisions) out of the sweep loops keeps the
// hot loops' size and register allocation what they are without them.
#ifndef JXF_NOINLINE
#ifdef __CUDACC__
#define JXF_NOINLINE __noinline__
#else
#define JXF_NOINLINE
#endif
#endif
struct Vec5 {
  double v[5];
};
template <int A>
__device__ JXF_NOINLINE Vec5 riemann_other(int variant, int sp, Vec5 PL, Vec5 PR, double gamma);   // defined below
// convective_solver = FLUX-SPLITTING (flux_splitting_scheme.py): bits 17-18 of the option word select the eigenvalue
// choice; only the STENCIL_GENERIC kernel instantiations carry the branch
enum { FS_NONE = 0, FS_ROE = 1, FS_CLLF = 2, FS_LLF = 3 };
struct Win6 {
  double w[5][6];
};
template <int A>
__device__ JXF_NOINLINE Vec5 flux_splitting_flux(Win6 W, double gamma, int id, int fs, int roe);    // defined below
// reconstruction_variable CONSERVATIVE / CHAR-CONSERVATIVE and frozen_state ROE: bits 19-21 of the option word
// (`mode` = variable | roe << 2), generic instantiations only
enum { VAR_PRIMITIVE = 0, VAR_CHAR_PRIMITIVE = 1, VAR_CONSERVATIVE = 2, VAR_CHAR_CONSERVATIVE = 3 };
struct Vec10 {
  double l[5], r[5];
};
template <int A>
__device__ JXF_NOINLINE Vec10 reconstruct_conservative(Win6 W, double gamma, int id, int mode);      // defined below
static __device__ JXF_NOINLINE double2 simple_signal_speeds(int sig, double uL, double uR, double aL, double aR, double rhoL,
                                                    double rhoR, double pL, double pR, double gamma) {
  double S_L, S_R;
  if (sig == SIG_ARITHMETIC) {
    const double u_mean = 0.5 * (uL + uR), a_mean = 0.5 * (aL + aR);
    S_L = fmin(u_mean - a_mean, uL - aL);
    S_R = fmax(u_mean + a_mean, uR + aR);
  } else if (sig == SIG_RUSANOV) {
    const double S_plus = fmax(fabs(uL) + aL, fabs(uR) + aR);
    S_L = -S_plus;
    S_R = S_plus;
  } else if (sig == SIG_DAVIS) {
    S_L = fmin(uL - aL, uR - aR);
    S_R = fmax(uL + aL, uR + aR);
  } else {   // SIG_TORO
    const double rho_bar = 0.5 * (rhoL + rhoR), a_bar = 0.5 * (aL + aR);
    const double p_pvrs = 0.5 * (pL + pR) - 0.5 * (uR - uL) * rho_bar * a_bar;
    const double p_star = fmax(0.0, p_pvrs);
    const double g_ = (gamma + 1) * 0.5 / gamma;
    const double qL = (p_star <= pL) ? 1.0 : sqrt(1 + g_ * (p_star / pL - 1));
    const double qR = (p_star <= pR) ? 1.0 : sqrt(1 + g_ * (p_star / pR - 1));
    S_L = uL - aL * qL;
    S_R = uR + aR * qR;
  }
  double2 r;
  r.x = S_L;
  r.y = S_R;
  return r;
}

// signal_speeds.py:109-133 in the reference's order (IEEE sqrt / division) -- for the HLL solver, which is not a
// tuned path (the HLLC kernels carry their own fast evaluation of the same formula)
static __device__ JXF_NOINLINE double2 einfeldt_signal_speeds(double uL, double uR, double aL, double aR, double rhoL,
                                                      double rhoR) {
  const double sL = gsqrt(rhoL), sR = gsqrt(rhoR);
  const double one_dens = grcp(sL + sR);
  const double eta2 = 0.5 * sL * sR * one_dens * one_dens;
  const double u_bar = (sL * uL + sR * uR) * one_dens;
  const double du = uR - uL;
  const double d_bar = gsqrt((sL * aL * aL + sR * aR * aR) * one_dens + eta2 * (du * du));
  double2 r;
  r.x = fmin(u_bar - d_bar, uL - aL);
  r.y = fmax(u_bar + d_bar, uR + aR);
  return r;
}

// velocity_minor_axes, equation_information.py:110
template <int A> struct AxisIds;
template <> struct AxisIds<0> { static constexpr int un = 1, t0 = 2, t1 = 3; };
template <> struct AxisIds<1> { static constexpr int un = 2, t0 = 3, t1 = 1; };
template <> struct AxisIds<2> { static constexpr int un = 3, t0 = 1, t1 = 2; };

// ---------------------------------------------------------------------------
// Generic reconstruction stencils (STENCIL_GENERIC).  q0..q5 = the six cells around the face in upwind-biased
// order (the face lies between q2 and q3): cells i-2..i+3 for the left state (j = 0), their mirror i+3..i-2 for
// the right state (j = 1) -- stencils/spatial_stencil.py:45-113.  Reference order, IEEE division.
//   WENO1            weno/weno1_js.py:24-29
//   WENO3-JS / -Z / -N  weno3_base.py:33-49, weno/weno3_js.py:15-43, weno/weno3_z.py:23-43, weno/weno3_n.py:22-41
//   CENTRAL2         reconstruction/central/central_2.py:38-47 (the same mean on both sides)
//   TENO5            teno/teno5.py:32-71 (C = 1, q = 6, C_T = 1e-5, d = (0.05, 0.55, 0.40)), weno5_base.py:34-51
//   WENO6-CU         weno6_base.py:32-58, weno/weno6_cu.py:36-63 (C = 20)
//   TENO6            teno6_base.py:32-62, teno/teno6.py:42-73 (C = 1, q = 6, C_T = 1e-7, d = (.05, .45, .3, .2))
//   TENO5-A / TENO6-A teno/teno5_a.py:46-101, teno/teno6_a.py:42-140: the cut-off C_T adapts to the local smoothness
//                    (TENO5-A reads the WENO5 weights of its base class -- its own are stored under another name)
//   KOREN .. VANLEER muscl/muscl3.py:39-77 with stencils/limiter.py:6-22 (the two sides are not mirror images of
//                    one formula, hence `j`)
// ---------------------------------------------------------------------------
__device__ __forceinline__ double teno_a_eta(double a, double b, double eps_d) {
  return gdiv(fabs(2.0 * a * b) + eps_d, a * a + b * b + eps_d);
}
__device__ __forceinline__ double teno_a_ct(double eta, double Cr, double alpha_1, double alpha_2) {
  const double m = 1.0 - fmin(1.0, gdiv(eta, Cr));
  const double x = 1.0 - m, x2 = x * x;
  const double g = (x2 * x2) * (1.0 + 4.0 * m);              // jnp.power(1 - m, 4) * (1 + 4 m)
  const double beta_bar = ceil(alpha_1 - alpha_2 * (1.0 - g)) - 1.0;
  return gpow10_neg(beta_bar);
}
// `id` known at compile time after inlining (reconstruct_generic_id below): only that stencil's branch survives.
__device__ __forceinline__ double stencil_generic_inl(const int id, const int j, double q0, double q1, double q2, double q3,
                                                      double q4, double q5) {
  const double eps = kStencilEps;
  if (id == ALT_WENO1) return q2;
  if (id == ALT_CENTRAL2) return 0.5 * ((j == 0) ? (q2 + q3) : (q3 + q2));    // buffer[i] + buffer[i+1] on both sides
  if (id == ALT_WENO3JS || id == ALT_WENO3Z || id == ALT_WENO3N) {
    const double d0 = q2 - q1, d1 = q3 - q2;
    const double beta_0 = d0 * d0, beta_1 = d1 * d1;
    double alpha_0, alpha_1;
    if (id == ALT_WENO3JS) {
      alpha_0 = (1.0 / 3.0) * grcp(beta_0 * beta_0 + eps);
      alpha_1 = (2.0 / 3.0) * grcp(beta_1 * beta_1 + eps);
    } else {
      double tau_3 = fabs(beta_0 - beta_1);
      if (id == ALT_WENO3N) {
        const double s = q1 - 2.0 * q2 + q3, t = q1 - q3;
        const double beta_3 = (13.0 / 12.0) * (s * s) + 0.25 * (t * t);
        tau_3 = fabs(0.5 * (beta_0 + beta_1) - beta_3);
      }
      alpha_0 = (1.0 / 3.0) * (1.0 + gdiv(tau_3, beta_0 + eps));
      alpha_1 = (2.0 / 3.0) * (1.0 + gdiv(tau_3, beta_1 + eps));
    }
    const double one_alpha = grcp(alpha_0 + alpha_1);
    const double p_0 = -0.5 * q1 + 1.5 * q2;
    const double p_1 = 0.5 * q2 + 0.5 * q3;
    return (alpha_0 * one_alpha) * p_0 + (alpha_1 * one_alpha) * p_1;
  }
  if (id == ALT_TENO5 || id == ALT_WENO6CU || id == ALT_TENO6 || id == ALT_WENO5Z || id == ALT_WENO5JS || id == ALT_TENO5A ||
      id == ALT_TENO6A) {
    const double s0 = q0 - 2.0 * q1 + q2, t0 = q0 - 4.0 * q1 + 3.0 * q2;
    const double s1 = q1 - 2.0 * q2 + q3, t1 = q1 - q3;
    const double s2 = q2 - 2.0 * q3 + q4, t2 = 3.0 * q2 - 4.0 * q3 + q4;
    const double beta_0 = (13.0 / 12.0) * (s0 * s0) + 0.25 * (t0 * t0);
    const double beta_1 = (13.0 / 12.0) * (s1 * s1) + 0.25 * (t1 * t1);
    const double beta_2 = (13.0 / 12.0) * (s2 * s2) + 0.25 * (t2 * t2);
    const double p_0 = (1.0 / 3.0) * q0 + (-7.0 / 6.0) * q1 + (11.0 / 6.0) * q2;
    const double p_1 = (-1.0 / 6.0) * q1 + (5.0 / 6.0) * q2 + (1.0 / 3.0) * q3;
    const double p_2 = (1.0 / 3.0) * q2 + (5.0 / 6.0) * q3 + (-1.0 / 6.0) * q4;
    if (id == ALT_WENO5Z || id == ALT_WENO5JS) {   // weno/weno5_z.py:32-52, weno/weno5_js.py:32-50
      double alpha_0, alpha_1, alpha_2;
      if (id == ALT_WENO5Z) {
        const double tau_5 = fabs(beta_0 - beta_2);
        alpha_0 = 0.1 * (1.0 + gdiv(tau_5, beta_0 + eps));
        alpha_1 = 0.6 * (1.0 + gdiv(tau_5, beta_1 + eps));
        alpha_2 = 0.3 * (1.0 + gdiv(tau_5, beta_2 + eps));
      } else {
        alpha_0 = 0.1 * grcp(beta_0 * beta_0 + eps);
        alpha_1 = 0.6 * grcp(beta_1 * beta_1 + eps);
        alpha_2 = 0.3 * grcp(beta_2 * beta_2 + eps);
      }
      const double one_alpha = grcp(alpha_0 + alpha_1 + alpha_2);
      return (alpha_0 * one_alpha) * p_0 + (alpha_1 * one_alpha) * p_1 + (alpha_2 * one_alpha) * p_2;
    }
    if (id == ALT_TENO5 || id == ALT_TENO5A) {
      const double tau_5 = fabs(beta_0 - beta_2);
      // jnp.power(x, 6): the value only feeds the cut-off comparison below
      const double x0 = 1.0 + gdiv(tau_5, beta_0 + eps), x1 = 1.0 + gdiv(tau_5, beta_1 + eps), x2 = 1.0 + gdiv(tau_5, beta_2 + eps);
      const double c0 = x0 * x0 * x0, c1 = x1 * x1 * x1, c2 = x2 * x2 * x2;
      const double gamma_0 = c0 * c0, gamma_1 = c1 * c1, gamma_2 = c2 * c2;
      const double one_gamma_sum = grcp(gamma_0 + gamma_1 + gamma_2);
      double CT = 1e-5, d0 = 0.05, d1 = 0.55, d2 = 0.40;
      if (id == ALT_TENO5A) {
        const double eps_d = 2.842105263157895e-07;          // 0.9 Cr / (1 - Cr) xi^2, Cr = 0.24, xi = 1e-3
        const double eta = fmin(teno_a_eta(q2 - q1, q1 - q0, eps_d),
                                fmin(teno_a_eta(q3 - q2, q2 - q1, eps_d), teno_a_eta(q4 - q3, q3 - q2, eps_d)));
        CT = teno_a_ct(eta, 0.24, 10.0, 5.0);
        d0 = 0.1; d1 = 0.6; d2 = 0.3;
      }
      const double w0 = d0 * ((gamma_0 * one_gamma_sum < CT) ? 0.0 : 1.0);
      const double w1 = d1 * ((gamma_1 * one_gamma_sum < CT) ? 0.0 : 1.0);
      const double w2 = d2 * ((gamma_2 * one_gamma_sum < CT) ? 0.0 : 1.0);
      const double one_dk = grcp(w0 + w1 + w2 + eps);
      return (w0 * one_dk) * p_0 + (w1 * one_dk) * p_1 + (w2 * one_dk) * p_2;
    }
    const double beta_6 = 1.0 / 10080 / 12 * (
        271779 * q0 * q0 +
        q0 * (-2380800 * q1 + 4086352 * q2 - 3462252 * q3 + 1458762 * q4 - 245620 * q5) +
        q1 * (5653317 * q1 - 20427884 * q2 + 17905032 * q3 - 7727988 * q4 + 1325006 * q5) +
        q2 * (19510972 * q2 - 35817664 * q3 + 15929912 * q4 - 2792660 * q5) +
        q3 * (17195652 * q3 - 15880404 * q4 + 2863984 * q5) +
        q4 * (3824847 * q4 - 1429976 * q5) +
        139633 * q5 * q5);
    if (id == ALT_TENO6 || id == ALT_TENO6A) {
      const bool adaptive = (id == ALT_TENO6A);
      double beta_6a = beta_6;
      double beta_3 = 1.0 / 240.0 * (
          q2 * (2107 * q2 - 9402 * q3 + 7042 * q4 - 1854 * q5)
          + q3 * (11003 * q3 - 17246 * q4 + 4642 * q5)
          + q4 * (7043 * q4 - 3882 * q5)
          + 547 * q5 * q5);
      const double p_3 = (3.0 / 12.0) * q2 + (13.0 / 12.0) * q3 + (-5.0 / 12.0) * q4 + (1.0 / 12.0) * q5;
      if (adaptive) {                                        // is_positivity_limiter_smoothness
        beta_3 = fabs(beta_3);
        beta_6a = fabs(beta_6a);
      }
      const double tau_6 = fabs(beta_6a - (1.0 / 6.0) * (beta_0 + 4.0 * beta_1 + beta_2));
      const double x0 = 1.0 + gdiv(tau_6, beta_0 + eps), x1 = 1.0 + gdiv(tau_6, beta_1 + eps);
      const double x2 = 1.0 + gdiv(tau_6, beta_2 + eps), x3 = 1.0 + gdiv(tau_6, beta_3 + eps);
      const double c0 = x0 * x0 * x0, c1 = x1 * x1 * x1, c2 = x2 * x2 * x2, c3 = x3 * x3 * x3;
      const double gamma_0 = c0 * c0, gamma_1 = c1 * c1, gamma_2 = c2 * c2, gamma_3 = c3 * c3;
      const double one_gamma_sum = grcp(gamma_0 + gamma_1 + gamma_2 + gamma_3);
      double CT = 1e-7, d0 = 0.050, d1 = 0.450, d2 = 0.300, d3 = 0.200;
      if (adaptive) {
        const double eps_d = 1.8433734939759037e-07;         // 0.9 Cr / (1 - Cr) xi^2, Cr = 0.17, xi = 1e-3
        const double f0 = q1 - q0, f1 = q2 - q1, f2 = q3 - q2, f3 = q4 - q3, f4 = q5 - q4;
        double eta = teno_a_eta(f1, f0, eps_d);
        eta = fmin(eta, teno_a_eta(f2, f1, eps_d));
        eta = fmin(eta, teno_a_eta(f3, f2, eps_d));
        eta = fmin(eta, teno_a_eta(f4, f3, eps_d));
        CT = teno_a_ct(eta, 0.17, 10.5, 4.5);
        d0 = 0.0855682281039113; d1 = 0.4294317718960898; d2 = 0.1727270875843552; d3 = 0.3122729124156450;
      }
      const double w0 = d0 * ((gamma_0 * one_gamma_sum < CT) ? 0.0 : 1.0);
      const double w1 = d1 * ((gamma_1 * one_gamma_sum < CT) ? 0.0 : 1.0);
      const double w2 = d2 * ((gamma_2 * one_gamma_sum < CT) ? 0.0 : 1.0);
      const double w3 = d3 * ((gamma_3 * one_gamma_sum < CT) ? 0.0 : 1.0);
      const double one_dk = adaptive ? grcp(w0 + w1 + w2 + w3) : grcp(w0 + w1 + w2 + w3 + eps);
      return (w0 * one_dk) * p_0 + (w1 * one_dk) * p_1 + (w2 * one_dk) * p_2 + (w3 * one_dk) * p_3;
    }
    const double beta_3 = beta_6;       // weno6_base.py calls the six-point indicator beta_3
    const double p_3 = (11.0 / 6.0) * q3 + (-7.0 / 6.0) * q4 + (1.0 / 3.0) * q5;
    const double tau_6 = beta_3 - (1.0 / 6.0) * (beta_0 + 4.0 * beta_1 + beta_2);
    const double alpha_0 = (1.0 / 20.0) * (20.0 + gdiv(tau_6, beta_0 + eps));
    const double alpha_1 = (9.0 / 20.0) * (20.0 + gdiv(tau_6, beta_1 + eps));
    const double alpha_2 = (9.0 / 20.0) * (20.0 + gdiv(tau_6, beta_2 + eps));
    const double alpha_3 = (1.0 / 20.0) * (20.0 + gdiv(tau_6, beta_3 + eps));
    const double one_alpha = grcp(alpha_0 + alpha_1 + alpha_2 + alpha_3);
    return (alpha_0 * one_alpha) * p_0 + (alpha_1 * one_alpha) * p_1 + (alpha_2 * one_alpha) * p_2 +
           (alpha_3 * one_alpha) * p_3;
  }
  // MUSCL3 with a slope limiter
  const double delta_central = (j == 0) ? (q3 - q2) : (q2 - q3);
  const double delta_upwind = (j == 0) ? (q2 - q1) : (q1 - q2);
  // one reciprocal: the two forms of the ratio differ in numerator / denominator only (select first, divide once)
  const bool big = delta_upwind >= eps;
  const double r = gdiv(big ? delta_central : delta_central + eps, big ? delta_upwind + 1e-10 : delta_upwind + eps);
  double lim;
  if (id == ALT_KOREN) lim = fmax(0.0, fmin(2.0 * r, fmin(gdiv(1.0 + 2.0 * r, 3.0), 2.0)));
  else if (id == ALT_MC) lim = fmax(0.0, fmin(2.0 * r, fmin((1.0 + r) / 2.0, 2.0)));
  else if (id == ALT_MINMOD) lim = fmax(0.0, fmin(1.0, r));
  else if (id == ALT_SUPERBEE) lim = fmax(0.0, fmax(fmin(1.0, 2.0 * r), fmin(2.0, r)));
  else if (id == ALT_VANALBADA) lim = gdiv(fmax(0.0, r) * (1.0 + r), 1.0 + r * r);
  else lim = gdiv(fmax(0.0, 2.0 * r), 1.0 + fabs(r));   // ALT_VANLEER
  return (j == 0) ? q2 + 0.5 * lim * delta_upwind : q2 - 0.5 * lim * delta_upwind;
}

// run-time id, out of line: the conservative-variable and flux-splitting paths
static __device__ JXF_NOINLINE double stencil_generic(int id, int j, double q0, double q1, double q2, double q3, double q4,
                                               double q5) {
  return stencil_generic_inl(id, j, q0, q1, q2, q3, q4, q5);
}
__device__ __forceinline__ void stencil_generic_lr(int id, const double (&q)[6], double& left, double& right) {
  left = stencil_generic(id, 0, q[0], q[1], q[2], q[3], q[4], q[5]);
  right = stencil_generic(id, 1, q[5], q[4], q[3], q[2], q[1], q[0]);
}
template <int ID>
__device__ __forceinline__ void stencil_const_lr(const double (&q)[6], double& left, double& right) {
  left = stencil_generic_inl(ID, 0, q[0], q[1], q[2], q[3], q[4], q[5]);
  right = stencil_generic_inl(ID, 1, q[5], q[4], q[3], q[2], q[1], q[0]);
}

// Frozen state of a face from the primitives of its two cells (eigendecomposition.py:120-281, single phase, ideal
// gas): ARITHMETIC :146-231 or ROE :233-276 (compute_roe_cons :283-294).  Reference order.
struct Frozen {
  double ave[5], H, G, c, cc, q2;
};
__device__ __forceinline__ double total_enthalpy_ref(const double (&p)[5], double gamma) {   // ideal_gas.py:90-110
  const double E = p[4] / (gamma - 1.0) + 0.5 * p[0] * ((p[1] * p[1] + p[2] * p[2]) + p[3] * p[3]);
  return gdiv(E + p[4], p[0]);
}
__device__ __forceinline__ Frozen frozen_state(const double (&pL)[5], const double (&pR)[5], double gamma, int roe) {
  Frozen f;
  if (!roe) {
#pragma unroll
    for (int v = 0; v < 5; ++v) f.ave[v] = 0.5 * (pL[v] + pR[v]);
    f.G = gamma - 1.0;
    f.H = total_enthalpy_ref(f.ave, gamma);
    f.c = gsqrt(gdiv(gamma * f.ave[4], f.ave[0]));
    f.cc = f.c * f.c;
    f.q2 = (f.ave[1] * f.ave[1] + f.ave[2] * f.ave[2]) + f.ave[3] * f.ave[3];
  } else {
    const double sL = gsqrt(pL[0]), sR = gsqrt(pR[0]);
#pragma unroll
    for (int v = 0; v < 5; ++v) f.ave[v] = gdiv(sL * pL[v] + sR * pR[v], sL + sR);
    f.ave[0] = gsqrt(pL[0] * pR[0]);
    const double rho_div = grcp(sL + sR);
    f.H = (sL * total_enthalpy_ref(pL, gamma) + sR * total_enthalpy_ref(pR, gamma)) * rho_div;
    const double psi = (sL * gdiv(pL[4], pL[0]) + sR * gdiv(pR[4], pR[0])) * rho_div;
    f.G = (sL * (gamma - 1.0) + sR * (gamma - 1.0)) * rho_div;
    const double du = pR[1] - pL[1], dv = pR[2] - pL[2], dw = pR[3] - pL[3];
    const double dq2 = (du * du + dv * dv) + dw * dw;
    const double p_over_rho = (gdiv(sL * pL[4], pL[0]) + gdiv(sR * pR[4], pR[0])) * rho_div + 0.5 * f.ave[0] * rho_div * rho_div * dq2;
    f.q2 = (f.ave[1] * f.ave[1] + f.ave[2] * f.ave[2]) + f.ave[3] * f.ave[3];
    f.cc = psi + f.G * p_over_rho;
    f.c = gsqrt(f.cc);
  }
  return f;
}

// reconstruct() of the generic stencils: PRIMITIVE (high_order_godunov.py:267-280), CHAR-PRIMITIVE (:298-316 with
// eigendecomposition.py:120-281, 425-431, 517-521) -- or, out of line, the two conservative forms -- reference order.
// `mode` = reconstruction variable | (frozen_state == ROE) << 2.
// One face, one stencil: the window comes by reference, the ten stencil evaluations (5 variables x 2 sides) are inlined
// with the stencil id a compile-time constant -- ONE out-of-line call per face instead of ten calls that each walk the
// id chain and spill / reload around the call (the rows kernel spent 55 % of its stall samples there,
// profiles/stalls_r02t_teno6a_rows.txt).  Same operations in the same order as before.
template <int A, bool CHAR, int ID>
static __device__ JXF_NOINLINE void reconstruct_generic_id(const double (&w)[5][6], double gamma, double (&pl)[5],
                                                           double (&pr)[5], int mode) {
  using Id = AxisIds<A>;
  // the TENO6 / adaptive stencils are 500-700 instructions per evaluation: their five rows are a rolled loop over a staged
  // copy of the rows, two evaluations per iteration (measured at 256^3, profiles/r02x_ / r02y_generic_timing.txt: TENO6-A
  // z + epilogue 7.09 ms with ten inlined copies, 5.80 ms rolled; WENO6-CU the other way round, 786 vs 667 MCUPS)
  constexpr bool kRolled = (ID == ALT_TENO6 || ID == ALT_TENO6A || ID == ALT_TENO5A);
  if (!CHAR) {
    if constexpr (kRolled) {
#pragma unroll 1
      for (int v = 0; v < 5; ++v) stencil_const_lr<ID>(w[v], pl[v], pr[v]);
    } else {
#pragma unroll
      for (int v = 0; v < 5; ++v) stencil_const_lr<ID>(w[v], pl[v], pr[v]);
    }
  } else {
    double cL[5], cR[5];
#pragma unroll
    for (int v = 0; v < 5; ++v) { cL[v] = w[v][2]; cR[v] = w[v][3]; }
    const Frozen fz = frozen_state(cL, cR, gamma, (mode >> 2) & 1);
    const double rho_ave = fz.ave[0];
    const double c_ave = fz.c;
    const double cc_ave = fz.cc;
    const double k_u = gdiv(0.5, c_ave);
    const double k_p = gdiv(0.5, cc_ave * rho_ave);
    const double k_cc = grcp(cc_ave);
    double q[6], l0, r0, l1, r1, l4, r4;
    if constexpr (kRolled) {
      // rows 0..2: the three characteristic fields, rows 3, 4: the transverse velocities
      double rows[5][6], L[5], R[5];
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        rows[0][k] = -k_u * w[Id::un][k] + k_p * w[4][k];
        rows[1][k] = w[0][k] - k_cc * w[4][k];
        rows[2][k] = k_u * w[Id::un][k] + k_p * w[4][k];
        rows[3][k] = w[Id::t0][k];
        rows[4][k] = w[Id::t1][k];
      }
#pragma unroll 1
      for (int v = 0; v < 5; ++v) stencil_const_lr<ID>(rows[v], L[v], R[v]);
      l0 = L[0]; r0 = R[0]; l1 = L[1]; r1 = R[1]; l4 = L[2]; r4 = R[2];
      pl[Id::t0] = L[3]; pr[Id::t0] = R[3];
      pl[Id::t1] = L[4]; pr[Id::t1] = R[4];
    } else {
#pragma unroll
      for (int k = 0; k < 6; ++k) q[k] = -k_u * w[Id::un][k] + k_p * w[4][k];
      stencil_const_lr<ID>(q, l0, r0);
#pragma unroll
      for (int k = 0; k < 6; ++k) q[k] = w[0][k] - k_cc * w[4][k];
      stencil_const_lr<ID>(q, l1, r1);
#pragma unroll
      for (int k = 0; k < 6; ++k) q[k] = k_u * w[Id::un][k] + k_p * w[4][k];
      stencil_const_lr<ID>(q, l4, r4);
      stencil_const_lr<ID>(w[Id::t0], pl[Id::t0], pr[Id::t0]);
      stencil_const_lr<ID>(w[Id::t1], pl[Id::t1], pr[Id::t1]);
    }
    const double ccr = cc_ave * rho_ave;
    pl[0] = rho_ave * (l0 + l4) + l1;
    pl[Id::un] = c_ave * (-l0 + l4);
    pl[4] = ccr * (l0 + l4);
    pr[0] = rho_ave * (r0 + r4) + r1;
    pr[Id::un] = c_ave * (-r0 + r4);
    pr[4] = ccr * (r0 + r4);
  }
}

template <int A, bool CHAR>
__device__ __forceinline__ void reconstruct_generic(const double (&w)[5][6], double gamma, double (&pl)[5],
                                                    double (&pr)[5], int id, int mode) {
  if ((mode & 3) >= VAR_CONSERVATIVE) {
    Win6 W;
#pragma unroll
    for (int v = 0; v < 5; ++v)
#pragma unroll
      for (int k = 0; k < 6; ++k) W.w[v][k] = w[v][k];
    const Vec10 o = reconstruct_conservative<A>(W, gamma, id, mode);
#pragma unroll
    for (int v = 0; v < 5; ++v) { pl[v] = o.l[v]; pr[v] = o.r[v]; }
    return;
  }
#define JXF_STENCIL_CASE(ID) case ID: reconstruct_generic_id<A, CHAR, ID>(w, gamma, pl, pr, mode); break;
  switch (id) {
    JXF_STENCIL_CASE(ALT_WENO5Z) JXF_STENCIL_CASE(ALT_WENO5JS) JXF_STENCIL_CASE(ALT_WENO1) JXF_STENCIL_CASE(ALT_WENO3JS)
    JXF_STENCIL_CASE(ALT_WENO3Z) JXF_STENCIL_CASE(ALT_TENO5) JXF_STENCIL_CASE(ALT_WENO6CU) JXF_STENCIL_CASE(ALT_KOREN)
    JXF_STENCIL_CASE(ALT_MC) JXF_STENCIL_CASE(ALT_MINMOD) JXF_STENCIL_CASE(ALT_SUPERBEE) JXF_STENCIL_CASE(ALT_VANALBADA)
    JXF_STENCIL_CASE(ALT_VANLEER) JXF_STENCIL_CASE(ALT_WENO3N) JXF_STENCIL_CASE(ALT_CENTRAL2) JXF_STENCIL_CASE(ALT_TENO6)
    JXF_STENCIL_CASE(ALT_TENO5A) JXF_STENCIL_CASE(ALT_TENO6A)
    default: break;
  }
#undef JXF_STENCIL_CASE
}

// ===========================================================================
// Two evaluations of the same formulas:
//   JXF_REFERENCE_ORDER : the reference's operations in the reference's order (IEEE div/sqrt).
//                         Without FMA contraction this is bit-identical to the reference
//                         (tests/test_hostsim.py); it documents WHAT is computed.
//   default             : the production evaluation -- same formulas, re-associated so that
//                         one face costs ~680 FP64-pipe instructions instead of ~1900:
//                         * WENO5-Z on first differences, left+right share differences and
//                           squares; the three nonlinear weights share ONE reciprocal
//                           (omega_k = n_k / sum n with n_k = d_k (b_k+tau) prod_{j!=k} b_j,
//                           b_k = beta_k + eps), betas carried with a common factor 12/13;
//                         * the characteristic projection acts on the differences and the
//                           face value is cell value + back-projected correction (R L = I);
//                         * reciprocals / rsqrt by MUFU seed + Newton (no IEEE slow path),
//                           1/rho, sqrt(rho), a^2 = gamma p / rho shared inside HLLC;
//                         * only the HLLC star flux selected by sign(S*) is evaluated.
//                         Deviation from the reference order is O(1e-15) relative
//                         (tests/test_hostsim.py, tests/test_gpu_parity.py: <= 1e-12).
// ===========================================================================
#ifdef JXF_REFERENCE_ORDER

// ---------------------------------------------------------------------------
// WENO5-Z  (stencils/reconstruction/shock_capturing/weno5_base.py:34-51,
//           weno/weno5_z.py:32-52).  (a,b,c,d,e) = cells i-2..i+2 for the left
//           state, mirrored (i+3..i-1) for the right state.
// ---------------------------------------------------------------------------
template <int ST>
__device__ __forceinline__ double weno5(double a, double b, double c, double d, double e) {
  const double s0 = a - 2.0 * b + c;
  const double q0 = a - 4.0 * b + 3.0 * c;
  const double s1 = b - 2.0 * c + d;
  const double q1 = b - d;
  const double s2 = c - 2.0 * d + e;
  const double q2 = 3.0 * c - 4.0 * d + e;
  const double beta0 = (13.0 / 12.0) * (s0 * s0) + 0.25 * (q0 * q0);
  const double beta1 = (13.0 / 12.0) * (s1 * s1) + 0.25 * (q1 * q1);
  const double beta2 = (13.0 / 12.0) * (s2 * s2) + 0.25 * (q2 * q2);
  double alpha0, alpha1, alpha2;
  if (ST == STENCIL_WENO5Z) {
    const double tau5 = fabs(beta0 - beta2);
    alpha0 = 0.1 * (1.0 + tau5 / (beta0 + kStencilEps));
    alpha1 = 0.6 * (1.0 + tau5 / (beta1 + kStencilEps));
    alpha2 = 0.3 * (1.0 + tau5 / (beta2 + kStencilEps));
  } else {   // WENO5-JS, weno/weno5_js.py:36-43
    alpha0 = 0.1 * (1.0 / (beta0 * beta0 + kStencilEps));
    alpha1 = 0.6 * (1.0 / (beta1 * beta1 + kStencilEps));
    alpha2 = 0.3 * (1.0 / (beta2 * beta2 + kStencilEps));
  }
  const double inv = 1.0 / (alpha0 + alpha1 + alpha2);
  const double p0 = (1.0 / 3.0) * a + (-7.0 / 6.0) * b + (11.0 / 6.0) * c;
  const double p1 = (-1.0 / 6.0) * b + (5.0 / 6.0) * c + (1.0 / 3.0) * d;
  const double p2 = (1.0 / 3.0) * c + (5.0 / 6.0) * d + (-1.0 / 6.0) * e;
  return (alpha0 * inv) * p0 + (alpha1 * inv) * p1 + (alpha2 * inv) * p2;
}

template <int ST>
__device__ __forceinline__ void weno5z_lr(const double (&q)[6], double& left, double& right) {
  left = weno5<ST>(q[0], q[1], q[2], q[3], q[4]);
  right = weno5<ST>(q[5], q[4], q[3], q[2], q[1]);
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
                                            double (&pl)[5], double (&pr)[5], int alt = 0, int mode = 0) {
  using Id = AxisIds<A>;
  if constexpr ((RECON >> 1) == STENCIL_GENERIC) {
    reconstruct_generic<A, (RECON & 1) != RECON_PRIMITIVE>(w, gamma, pl, pr, alt, mode);
  } else if ((RECON & 1) == RECON_PRIMITIVE) {
#pragma unroll
    for (int v = 0; v < 5; ++v) weno5z_lr<(RECON >> 1)>(w[v], pl[v], pr[v]);
  } else {
    const double rho_ave = 0.5 * (w[0][2] + w[0][3]);
    const double p_ave = 0.5 * (w[4][2] + w[4][3]);
    const double c_ave = sqrt(gamma * p_ave / rho_ave);
    const double cc_ave = c_ave * c_ave;
    const double k_u = gdiv(0.5, c_ave);
    const double k_p = gdiv(0.5, cc_ave * rho_ave);
    const double k_cc = grcp(cc_ave);
    double q[6], l0, r0, l1, r1, l4, r4;
#pragma unroll
    for (int k = 0; k < 6; ++k) q[k] = -k_u * w[Id::un][k] + k_p * w[4][k];
    weno5z_lr<(RECON >> 1)>(q, l0, r0);
#pragma unroll
    for (int k = 0; k < 6; ++k) q[k] = w[0][k] - k_cc * w[4][k];
    weno5z_lr<(RECON >> 1)>(q, l1, r1);
#pragma unroll
    for (int k = 0; k < 6; ++k) q[k] = k_u * w[Id::un][k] + k_p * w[4][k];
    weno5z_lr<(RECON >> 1)>(q, l4, r4);
    weno5z_lr<(RECON >> 1)>(w[Id::t0], pl[Id::t0], pr[Id::t0]);
    weno5z_lr<(RECON >> 1)>(w[Id::t1], pl[Id::t1], pr[Id::t1]);
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
                                             double gamma, double (&F)[5], int sig = SIG_EINFELDT) {
  using Id = AxisIds<A>;
  double cl[5], cr[5];
  cons_from_prims(pl, gamma, cl);
  cons_from_prims(pr, gamma, cr);
  const double aL = sqrt(gamma * pl[4] / pl[0]);
  const double aR = sqrt(gamma * pr[4] / pr[0]);
  const double uL = pl[Id::un], uR = pr[Id::un];
  if (RIEMANN == RIEMANN_HLLC) {
    double S_L, S_R;
    if ((sig & 15) == SIG_EINFELDT) {
      const double sL = sqrt(pl[0]), sR = sqrt(pr[0]);
      const double one_dens = 1.0 / (sL + sR);
      const double eta2 = 0.5 * sL * sR * one_dens * one_dens;
      const double u_bar = (sL * uL + sR * uR) * one_dens;
      const double du = uR - uL;
      const double d_bar = sqrt((sL * aL * aL + sR * aR * aR) * one_dens + eta2 * (du * du));
      S_L = fmin(u_bar - d_bar, uL - aL);
      S_R = fmax(u_bar + d_bar, uR + aR);
    } else {
      const double2 ss = simple_signal_speeds(sig & 15, uL, uR, aL, aR, pl[0], pr[0], pl[4], pr[4], gamma);
      S_L = ss.x;
      S_R = ss.y;
    }
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
  } else if ((sig >> 11) & 3) {
    Vec5 a, b;
#pragma unroll
    for (int v = 0; v < 5; ++v) { a.v[v] = pl[v]; b.v[v] = pr[v]; }
    const Vec5 o = riemann_other<A>((sig >> 11) & 3, sig & 15, a, b, gamma);
#pragma unroll
    for (int v = 0; v < 5; ++v) F[v] = o.v[v];
  } else if ((sig >> 4) & 1) {
    // HLL: solvers/riemann_solvers/HLL.py (the kernels' RIEMANN_RUSANOV instantiation with the HLL bit of `sig`)
    const int sp = sig & 15;
    const double2 ss = (sp == SIG_EINFELDT) ? einfeldt_signal_speeds(uL, uR, aL, aR, pl[0], pr[0])
                                            : simple_signal_speeds(sp, uL, uR, aL, aR, pl[0], pr[0], pl[4], pr[4], gamma);
    const double wL = fmin(ss.x, 0.0), wR = fmax(ss.y, 0.0);
    double fl[5], fr[5];
    physical_flux<A>(pl, cl, fl);
    physical_flux<A>(pr, cr, fr);
#pragma unroll
    for (int v = 0; v < 5; ++v)
      F[v] = (wR * fl[v] - wL * fr[v] + wL * wR * (cr[v] - cl[v])) / (wR - wL + kEps);
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

#else  // production evaluation

#ifndef JXF_HLLC_BRANCH
#define JXF_HLLC_BRANCH 1
#endif
#ifndef JXF_RIEMANN_MAIN       // marching sweeps: 1 = branch-free short-chain Riemann solve + rare S* = 0 fix-up
#define JXF_RIEMANN_MAIN 0
#endif

// ---------------------------------------------------------------------------
// WENO5-Z on the five first differences d_i = q_{i+1} - q_i of the 6-cell window
// (weno5_base.py:34-51, weno/weno5_z.py:32-52).  Returns the corrections
//   left  face value = q2 + cl   (cells i-2..i+2, centre i)
//   right face value = q3 + cr   (mirror cells i+3..i-1, centre i+1)
// beta~_k = (12/13) beta_k = e^2 + (3/13) t^2 ; tau, eps scale alike so tau/(beta+eps) is unchanged.
// ---------------------------------------------------------------------------
// WENO5-JS (weno/weno5_js.py:32-50): alpha_k = d_k / (beta_k^2 + eps); with B_k = beta~_k^2 + eps~^2 (the same
// common-factor scaling, (12/13)^2 on both terms) omega_k = d_k prod_{j!=k} B_j / sum -- the same one-reciprocal
// form with tau = 0 and b_k replaced by B_k.
template <int ST>
__device__ __forceinline__ void weno5_corr(double d0, double d1, double d2, double d3, double d4,
                                           double& cl, double& cr) {
  constexpr double k = 3.0 / 13.0;
  constexpr double eps = (ST == STENCIL_WENO5Z) ? kStencilEps * (12.0 / 13.0) : 0.0;
  constexpr double eps_js = kStencilEps * (12.0 / 13.0) * (12.0 / 13.0);
  const double e1 = d1 - d0, e2 = d2 - d1, e3 = d3 - d2, e4 = d4 - d3;      // second differences
  // eps is folded into the squared second differences: beta~_k + eps~ = k t^2 + (e^2 + eps~)
  const double s1 = fma(e1, e1, eps), s2 = fma(e2, e2, eps), s3 = fma(e3, e3, eps), s4 = fma(e4, e4, eps);
  // left stencils:  (a-4b+3c) = 3 d1 - d0 ; (b-d) = -(d1+d2) ; (3c-4d+e) = d3 - 3 d2
  const double tl0 = fma(3.0, d1, -d0), tl1 = d1 + d2, tl2 = fma(-3.0, d2, d3);
  // right stencils (mirrored): d4 - 3 d3 ; d2 + d3 ; 3 d2 - d1
  const double tr0 = fma(-3.0, d3, d4), tr1 = d2 + d3, tr2 = fma(3.0, d2, -d1);
  double bl0 = fma(k, tl0 * tl0, s1), bl1 = fma(k, tl1 * tl1, s2), bl2 = fma(k, tl2 * tl2, s3);
  double br0 = fma(k, tr0 * tr0, s4), br1 = fma(k, tr1 * tr1, s3), br2 = fma(k, tr2 * tr2, s2);
  if (ST == STENCIL_WENO5JS) {
    bl0 = fma(bl0, bl0, eps_js); bl1 = fma(bl1, bl1, eps_js); bl2 = fma(bl2, bl2, eps_js);
    br0 = fma(br0, br0, eps_js); br1 = fma(br1, br1, eps_js); br2 = fma(br2, br2, eps_js);
  }
  {
    const double tau = (ST == STENCIL_WENO5Z) ? fabs(bl0 - bl2) : 0.0;           // eps cancels in the difference
    // n_k = d_k (b_k + tau) prod_{j != k} b_j = d_k (P + tau prod_{j != k} b_j), P = b0 b1 b2, with the
    // common factor 1/10 dropped: d = (1, 6, 3)
    const double p12 = bl1 * bl2, p02 = bl0 * bl2, p01 = bl0 * bl1;
    const double P = bl0 * p12;
    // Z: (b_k + tau) prod_{j != k} b_j ; JS: prod_{j != k} B_j
    const double n0 = (ST == STENCIL_WENO5Z) ? fma(tau, p12, P) : p12;
    const double m1 = (ST == STENCIL_WENO5Z) ? fma(tau, p02, P) : p02;
    const double m2 = (ST == STENCIL_WENO5Z) ? fma(tau, p01, P) : p01;
    const double den = fma(6.0, m1, fma(3.0, m2, n0));
    // p_k - c:  p0-c = 5/6 d1 - 1/3 d0 ; p1-c = 1/3 d2 + 1/6 d1 ; p2-c = 2/3 d2 - 1/6 d3   (x d_k)
    const double q0 = fma(5.0 / 6.0, d1, (-1.0 / 3.0) * d0);
    const double q1 = fma(2.0, d2, d1);                          // 6 (p1-c)
    const double q2 = fma(2.0, d2, -0.5 * d3);                   // 3 (p2-c)
    const double num = fma(m2, q2, fma(m1, q1, n0 * q0));
    cl = num * rcp_fast(den);
  }
  {
    const double tau = (ST == STENCIL_WENO5Z) ? fabs(br0 - br2) : 0.0;
    const double p12 = br1 * br2, p02 = br0 * br2, p01 = br0 * br1;
    const double P = br0 * p12;
    const double n0 = (ST == STENCIL_WENO5Z) ? fma(tau, p12, P) : p12;
    const double m1 = (ST == STENCIL_WENO5Z) ? fma(tau, p02, P) : p02;
    const double m2 = (ST == STENCIL_WENO5Z) ? fma(tau, p01, P) : p01;
    const double den = fma(6.0, m1, fma(3.0, m2, n0));
    // mirrored differences d0'=-d4, d1'=-d3, d2'=-d2, d3'=-d1
    const double q0 = fma(-5.0 / 6.0, d3, (1.0 / 3.0) * d4);
    const double q1 = -fma(2.0, d2, d3);
    const double q2 = fma(-2.0, d2, 0.5 * d1);
    const double num = fma(m2, q2, fma(m1, q1, n0 * q0));
    cr = num * rcp_fast(den);
  }
}

// ---------------------------------------------------------------------------
// The same scheme split at the cell: the five cells (i-2..i+2) around cell i feed BOTH the left
// state of face i+1/2 and (mirrored) the right state of face i-1/2, with the same three smoothness
// indicators and tau (beta_k^R = beta_{2-k}^L).  weno5z_g evaluates the d-free weight products
// g_k = (b_k + tau) prod_{j!=k} b_j of that cell once from its four differences
// (D0..D3 = differences of cells i-2..i+2); the two face values then cost one reciprocal each.
// Used by the marching (strided) sweeps for fields that are reconstructed as they are (all five in
// PRIMITIVE mode, the two tangential velocities in CHAR-PRIMITIVE mode): the weights computed for the
// right state at one face are carried in registers to the left state of the next face.
// ---------------------------------------------------------------------------
struct WenoG {
  double g0, g1, g2;
};

template <int ST>
__device__ __forceinline__ WenoG weno5z_g(double D0, double D1, double D2, double D3) {
  constexpr double k = 3.0 / 13.0;
  constexpr double eps = (ST == STENCIL_WENO5Z) ? kStencilEps * (12.0 / 13.0) : 0.0;
  constexpr double eps_js = kStencilEps * (12.0 / 13.0) * (12.0 / 13.0);
  const double e1 = D1 - D0, e2 = D2 - D1, e3 = D3 - D2;
  const double s1 = fma(e1, e1, eps), s2 = fma(e2, e2, eps), s3 = fma(e3, e3, eps);
  const double t0 = fma(3.0, D1, -D0), t1 = D1 + D2, t2 = fma(-3.0, D2, D3);
  double b0 = fma(k, t0 * t0, s1), b1 = fma(k, t1 * t1, s2), b2 = fma(k, t2 * t2, s3);
  if (ST == STENCIL_WENO5JS) {
    b0 = fma(b0, b0, eps_js); b1 = fma(b1, b1, eps_js); b2 = fma(b2, b2, eps_js);
  }
  const double tau = (ST == STENCIL_WENO5Z) ? fabs(b0 - b2) : 0.0;
  const double p12 = b1 * b2, p02 = b0 * b2, p01 = b0 * b1;
  const double P = b0 * p12;
  WenoG g;
  g.g0 = (ST == STENCIL_WENO5Z) ? fma(tau, p12, P) : p12;
  g.g1 = (ST == STENCIL_WENO5Z) ? fma(tau, p02, P) : p02;
  g.g2 = (ST == STENCIL_WENO5Z) ? fma(tau, p01, P) : p01;
  return g;
}
// left state of the face to the right of the cell: cell value + this
__device__ __forceinline__ double weno5z_left_corr(const WenoG& g, double D0, double D1, double D2, double D3) {
  const double den = fma(6.0, g.g1, fma(3.0, g.g2, g.g0));
  const double q0 = fma(5.0 / 6.0, D1, (-1.0 / 3.0) * D0);
  const double q1 = fma(2.0, D2, D1);
  const double q2 = fma(2.0, D2, -0.5 * D3);
  return fma(g.g2, q2, fma(g.g1, q1, g.g0 * q0)) * rcp_fast(den);
}
// right state of the face to the left of the cell (mirror: sub-stencil k <-> 2-k): cell value + this
__device__ __forceinline__ double weno5z_right_corr(const WenoG& g, double D0, double D1, double D2, double D3) {
  const double den = fma(6.0, g.g1, fma(3.0, g.g0, g.g2));
  const double q0 = fma(-5.0 / 6.0, D2, (1.0 / 3.0) * D3);
  const double q1 = -fma(2.0, D1, D2);
  const double q2 = fma(-2.0, D1, 0.5 * D0);
  return fma(g.g0, q2, fma(g.g1, q1, g.g2 * q0)) * rcp_fast(den);
}

// ---------------------------------------------------------------------------
// EOS / variable transforms (ideal_gas.py:69-88, equation_manager.py:93-101, 164-171, 237-252)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void cons_from_prims(const double (&p)[5], double gamma, double (&c)[5]) {
  const double e = p[4] / (p[0] * (gamma - 1.0));
  c[0] = p[0];
  c[1] = p[0] * p[1];
  c[2] = p[0] * p[2];
  c[3] = p[0] * p[3];
  c[4] = p[0] * (0.5 * ((p[1] * p[1] + p[2] * p[2]) + p[3] * p[3]) + e);
}

// E = rho (u.u/2) + p/(gamma-1) with ig1 = 1/(gamma-1): no division
__device__ __forceinline__ void cons_from_prims_fast(const double (&p)[5], double ig1, double (&c)[5]) {
  const double q = fma(p[3], p[3], fma(p[2], p[2], p[1] * p[1]));
  c[0] = p[0];
  c[1] = p[0] * p[1];
  c[2] = p[0] * p[2];
  c[3] = p[0] * p[3];
  c[4] = fma(p[0], 0.5 * q, p[4] * ig1);
}

__device__ __forceinline__ void prims_from_cons(const double (&c)[5], double gamma, double (&p)[5]) {
  const double one_rho = rcp_fast(c[0]);
  p[0] = c[0];
  p[1] = c[1] * one_rho;
  p[2] = c[2] * one_rho;
  p[3] = c[3] * one_rho;
  const double e = c[4] * one_rho - 0.5 * ((p[1] * p[1] + p[2] * p[2]) + p[3] * p[3]);
  p[4] = (gamma - 1.0) * e * c[0];
}

// ---------------------------------------------------------------------------
// Reconstruction (high_order_godunov.py:267-280 PRIMITIVE, :298-316 CHAR-PRIMITIVE with
// eigendecomposition.py:139-148,215-231 frozen state, :425-431 projection, :517-521 back-projection)
// ---------------------------------------------------------------------------
template <int A, int RECON>
__device__ __forceinline__ void reconstruct(const double (&w)[5][6], double gamma,
                                            double (&pl)[5], double (&pr)[5], int alt = 0, int mode = 0) {
  using Id = AxisIds<A>;
  if constexpr ((RECON >> 1) == STENCIL_GENERIC) {
    reconstruct_generic<A, (RECON & 1) != RECON_PRIMITIVE>(w, gamma, pl, pr, alt, mode);
  } else if ((RECON & 1) == RECON_PRIMITIVE) {
#pragma unroll
    for (int v = 0; v < 5; ++v) {
      double cl, cr;
      weno5_corr<(RECON >> 1)>(w[v][1] - w[v][0], w[v][2] - w[v][1], w[v][3] - w[v][2], w[v][4] - w[v][3], w[v][5] - w[v][4], cl, cr);
      pl[v] = w[v][2] + cl;
      pr[v] = w[v][3] + cr;
    }
  } else {
    double dr[5], du[5], dp[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      dr[k] = w[0][k + 1] - w[0][k];
      du[k] = w[Id::un][k + 1] - w[Id::un][k];
      dp[k] = w[4][k + 1] - w[4][k];
    }
    // frozen state = arithmetic mean of cells i, i+1
    const double rho_ave = fma(0.5, dr[2], w[0][2]);
    const double p_ave = fma(0.5, dp[2], w[4][2]);
    const double gp = gamma * p_ave;                  // = cc_ave * rho_ave
    // z = 1/sqrt(gp rho): c = gp z, 1/c = rho z, 1/cc = (rho z)^2, 0.5/(cc rho) = 0.5 rho z^2
    const double z = rsqrt_fast(gp * rho_ave);
    const double ic = rho_ave * z;                    // 1 / c_ave
    const double c_ave = gp * z;
    const double k_u = 0.5 * ic;                      // 0.5 / c
    const double k_cc = ic * ic;                      // 1 / cc
    const double k_p = k_u * z;                       // 0.5 / (cc rho)
    double l0, r0, l1, r1, l4, r4;
    {
      double a[5], b[5], c[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const double t = k_p * dp[k];
        a[k] = fma(-k_u, du[k], t);                   // d W0
        c[k] = fma(k_u, du[k], t);                    // d W4
        b[k] = fma(-k_cc, dp[k], dr[k]);              // d W1
      }
      weno5_corr<(RECON >> 1)>(a[0], a[1], a[2], a[3], a[4], l0, r0);
      weno5_corr<(RECON >> 1)>(b[0], b[1], b[2], b[3], b[4], l1, r1);
      weno5_corr<(RECON >> 1)>(c[0], c[1], c[2], c[3], c[4], l4, r4);
    }
    double tl, tr;
    weno5_corr<(RECON >> 1)>(w[Id::t0][1] - w[Id::t0][0], w[Id::t0][2] - w[Id::t0][1], w[Id::t0][3] - w[Id::t0][2],
                w[Id::t0][4] - w[Id::t0][3], w[Id::t0][5] - w[Id::t0][4], tl, tr);
    pl[Id::t0] = w[Id::t0][2] + tl;
    pr[Id::t0] = w[Id::t0][3] + tr;
    weno5_corr<(RECON >> 1)>(w[Id::t1][1] - w[Id::t1][0], w[Id::t1][2] - w[Id::t1][1], w[Id::t1][3] - w[Id::t1][2],
                w[Id::t1][4] - w[Id::t1][3], w[Id::t1][5] - w[Id::t1][4], tl, tr);
    pl[Id::t1] = w[Id::t1][2] + tl;
    pr[Id::t1] = w[Id::t1][3] + tr;
    // prims = cell value + R * correction
    const double sl = l0 + l4, sr = r0 + r4;
    pl[0] = w[0][2] + fma(rho_ave, sl, l1);
    pl[Id::un] = fma(c_ave, l4 - l0, w[Id::un][2]);
    pl[4] = fma(gp, sl, w[4][2]);
    pr[0] = w[0][3] + fma(rho_ave, sr, r1);
    pr[Id::un] = fma(c_ave, r4 - r0, w[Id::un][3]);
    pr[4] = fma(gp, sr, w[4][3]);
  }
}

// Carry of the cell-centred weights between consecutive faces of a marching sweep.
template <int RECON>
struct ReconCarry {
  // fields reconstructed as they are (none for the generic stencils: nothing is carried)
  static constexpr int N = ((RECON >> 1) == STENCIL_GENERIC) ? 0 : (((RECON & 1) == RECON_PRIMITIVE) ? 5 : 2);
  WenoG g[N > 0 ? N : 1];
};

// weights of the cell that is the window's cell k=1..: prime the carry from the 5 cells w[.][0..4]
// (= the cell-centred set of window cell 2, i.e. the left stencil of this face)
template <int A, int RECON>
__device__ __forceinline__ void recon_carry_init(const double (&w)[5][6], ReconCarry<RECON>& cy) {
  using Id = AxisIds<A>;
  if constexpr ((RECON >> 1) != STENCIL_GENERIC) {
#pragma unroll
    for (int j = 0; j < ReconCarry<RECON>::N; ++j) {
      const int v = ((RECON & 1) == RECON_PRIMITIVE) ? j : (j == 0 ? Id::t0 : Id::t1);
      cy.g[j] = weno5z_g<(RECON >> 1)>(w[v][1] - w[v][0], w[v][2] - w[v][1], w[v][3] - w[v][2], w[v][4] - w[v][3]);
    }
  }
}

// weights of the face's RIGHT cell (window cell 3) for the as-is fields
template <int A, int RECON>
__device__ __forceinline__ void recon_g_right(const double (&w)[5][6], ReconCarry<RECON>& gr) {
  using Id = AxisIds<A>;
  if constexpr ((RECON >> 1) != STENCIL_GENERIC) {
#pragma unroll
    for (int j = 0; j < ReconCarry<RECON>::N; ++j) {
      const int v = ((RECON & 1) == RECON_PRIMITIVE) ? j : (j == 0 ? Id::t0 : Id::t1);
      gr.g[j] = weno5z_g<(RECON >> 1)>(w[v][2] - w[v][1], w[v][3] - w[v][2], w[v][4] - w[v][3], w[v][5] - w[v][4]);
    }
  }
}

// reconstruct() with the cell-centred weights of the as-is fields GIVEN: gl = those of window cell 2 (left stencil of
// this face), gr = those of window cell 3 (right stencil).  Marching sweeps carry gr of one face to gl of the next in
// registers (reconstruct_carry); sweeps whose lanes are consecutive faces (sweep_rows) get gl from the lane before by
// warp shuffle -- either way every cell's weights are evaluated once.  Tuned stencils only.
template <int A, int RECON>
__device__ __forceinline__ void reconstruct_given(const double (&w)[5][6], double gamma, double (&pl)[5], double (&pr)[5],
                                                  const ReconCarry<RECON>& gl, const ReconCarry<RECON>& gr) {
  using Id = AxisIds<A>;
#pragma unroll
  for (int j = 0; j < ReconCarry<RECON>::N; ++j) {
    const int v = ((RECON & 1) == RECON_PRIMITIVE) ? j : (j == 0 ? Id::t0 : Id::t1);
    const double d0 = w[v][1] - w[v][0], d1 = w[v][2] - w[v][1], d2 = w[v][3] - w[v][2], d3 = w[v][4] - w[v][3],
                 d4 = w[v][5] - w[v][4];
    pl[v] = w[v][2] + weno5z_left_corr(gl.g[j], d0, d1, d2, d3);
    pr[v] = w[v][3] + weno5z_right_corr(gr.g[j], d1, d2, d3, d4);
  }
  if ((RECON & 1) != RECON_PRIMITIVE) {
    double dr[5], du[5], dp[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      dr[k] = w[0][k + 1] - w[0][k];
      du[k] = w[Id::un][k + 1] - w[Id::un][k];
      dp[k] = w[4][k + 1] - w[4][k];
    }
    const double rho_ave = fma(0.5, dr[2], w[0][2]);
    const double p_ave = fma(0.5, dp[2], w[4][2]);
    const double gp = gamma * p_ave;
    const double z = rsqrt_fast(gp * rho_ave);
    const double ic = rho_ave * z;
    const double c_ave = gp * z;
    const double k_u = 0.5 * ic;
    const double k_cc = ic * ic;
    const double k_p = k_u * z;
    double l0, r0, l1, r1, l4, r4;
    {
      double a[5], b[5], c[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const double t = k_p * dp[k];
        a[k] = fma(-k_u, du[k], t);
        c[k] = fma(k_u, du[k], t);
        b[k] = fma(-k_cc, dp[k], dr[k]);
      }
      weno5_corr<(RECON >> 1)>(a[0], a[1], a[2], a[3], a[4], l0, r0);
      weno5_corr<(RECON >> 1)>(b[0], b[1], b[2], b[3], b[4], l1, r1);
      weno5_corr<(RECON >> 1)>(c[0], c[1], c[2], c[3], c[4], l4, r4);
    }
    const double sl = l0 + l4, sr = r0 + r4;
    pl[0] = w[0][2] + fma(rho_ave, sl, l1);
    pl[Id::un] = fma(c_ave, l4 - l0, w[Id::un][2]);
    pl[4] = fma(gp, sl, w[4][2]);
    pr[0] = w[0][3] + fma(rho_ave, sr, r1);
    pr[Id::un] = fma(c_ave, r4 - r0, w[Id::un][3]);
    pr[4] = fma(gp, sr, w[4][3]);
  }
}

// reconstruct() for marching sweeps: `cy` holds, on entry, the weights of window cell 2 (left stencil
// of this face); on exit those of window cell 3 (right stencil of this face = left stencil of the next)
template <int A, int RECON>
__device__ __forceinline__ void reconstruct_carry(const double (&w)[5][6], double gamma, double (&pl)[5],
                                                  double (&pr)[5], ReconCarry<RECON>& cy, int alt = 0, int mode = 0) {
  if constexpr ((RECON >> 1) == STENCIL_GENERIC) {
    reconstruct<A, RECON>(w, gamma, pl, pr, alt, mode);
  } else {
    ReconCarry<RECON> gn;
    recon_g_right<A, RECON>(w, gn);
    reconstruct_given<A, RECON>(w, gamma, pl, pr, cy, gn);
    cy = gn;
  }
}

// ---------------------------------------------------------------------------
// Riemann solvers: HLLC + Einfeldt (HLLC.py:41-126, signal_speeds.py:109-133, :159-199),
// Rusanov (Rusanov.py:25-47)
// ---------------------------------------------------------------------------
template <int A>
__device__ __forceinline__ void physical_flux(const double (&p)[5], const double (&c)[5], double (&f)[5]) {
  const double m = c[1 + A];
  f[0] = m;
  f[1] = m * p[1];
  f[2] = m * p[2];
  f[3] = m * p[3];
  f[1 + A] = fma(m, p[1 + A], p[4]);
  f[4] = p[1 + A] * (c[4] + p[4]);
}

// F*_K = F_K + S_K^{-/+} (U*_K - U_K), Toro 10.72/10.73; dK = rho_K (S_K - u_K).  The conservative
// state of side K is formed here, i.e. only for the side(s) that sign(S*) selects.
template <int A>
__device__ __forceinline__ void hllc_star_flux(const double (&p)[5], double ig1, double inv_rho, double S_K,
                                               double S_lim, double dK, double S_star, double (&fs)[5]) {
  using Id = AxisIds<A>;
  double c[5];
  cons_from_prims_fast(p, ig1, c);
  const double pre = dK * rcp_fast(S_K - S_star);                  // (S_K-u_K)/(S_K-S*) rho_K
  const double es = fma(S_star - p[Id::un], fma(p[4], rcp_fast(dK), S_star), c[4] * inv_rho);
  double us[5];
  us[0] = pre;
  us[Id::un] = pre * S_star;
  us[Id::t0] = pre * p[Id::t0];
  us[Id::t1] = pre * p[Id::t1];
  us[4] = pre * es;
  double f[5];
  physical_flux<A>(p, c, f);
#pragma unroll
  for (int v = 0; v < 5; ++v) fs[v] = fma(S_lim, us[v] - c[v], f[v]);
}

template <int A, int RIEMANN>
__device__ __forceinline__ void riemann_flux(const double (&pl)[5], const double (&pr)[5],
                                             double gamma, double (&F)[5], int sig = SIG_EINFELDT) {
  using Id = AxisIds<A>;
  const double ig1 = 1.0 / (gamma - 1.0);
  const double uL = pl[Id::un], uR = pr[Id::un];
  // y = 1/sqrt(rho): 1/rho = y^2, sqrt(rho) = rho y
  const double yL = rsqrt_fast(pl[0]), yR = rsqrt_fast(pr[0]);
  const double irL = yL * yL, irR = yR * yR;
  const double a2L = gamma * pl[4] * irL, a2R = gamma * pr[4] * irR;       // a^2
  const double aL = sqrt_fast(a2L, rsqrt_fast(a2L)), aR = sqrt_fast(a2R, rsqrt_fast(a2R));
  if (RIEMANN == RIEMANN_HLLC) {
    double S_L, S_R;
    if ((sig & 15) == SIG_EINFELDT) {
      const double sL = pl[0] * yL, sR = pr[0] * yR;                       // sqrt(rho)
      const double od = rcp_fast(sL + sR);
      const double eta2 = 0.5 * sL * sR * od * od;
      const double u_bar = fma(sL, uL, sR * uR) * od;
      const double du = uR - uL;
      const double x = fma(eta2, du * du, fma(sL, a2L, sR * a2R) * od);
      const double d_bar = sqrt_fast(x, rsqrt_fast(x));
      S_L = fmin(u_bar - d_bar, uL - aL);
      S_R = fmax(u_bar + d_bar, uR + aR);
    } else {      // the simple estimates (uniform branch; not the tuned path)
      const double2 ss = simple_signal_speeds(sig & 15, uL, uR, aL, aR, pl[0], pr[0], pl[4], pr[4], gamma);
      S_L = ss.x;
      S_R = ss.y;
    }
    const double dL = pl[0] * (S_L - uL);
    const double dR = pr[0] * (S_R - uR);
    const double S_star = ((pr[4] - pl[4]) + fma(uL, dL, -(uR * dR))) * rcp_fast(dL - dR);
    // F = 1/2 (1 + sign S*) F*_L + 1/2 (1 - sign S*) F*_R : only the selected side is evaluated
    // (S* > 0 -> F*_L, S* < 0 -> F*_R, S* = 0 -> the mean, sign(0) = 0 in the reference)
#if JXF_HLLC_BRANCH == 2
    // two inlined copies of the star flux instead of four (instruction-cache footprint of the hot loops):
    // S* >= 0 evaluates the left one, S* <= 0 the right one, S* = 0 both and their mean
#pragma unroll
    for (int v = 0; v < 5; ++v) F[v] = 0.0;
    if (S_star >= 0.0) hllc_star_flux<A>(pl, ig1, irL, S_L, fmin(S_L, 0.0), dL, S_star, F);
    if (S_star <= 0.0) {
      double fR[5];
      hllc_star_flux<A>(pr, ig1, irR, S_R, fmax(S_R, 0.0), dR, S_star, fR);
      const bool both = (S_star == 0.0);
#pragma unroll
      for (int v = 0; v < 5; ++v) F[v] = both ? 0.5 * (F[v] + fR[v]) : fR[v];
    }
#elif JXF_HLLC_BRANCH
    if (S_star > 0.0) {
      hllc_star_flux<A>(pl, ig1, irL, S_L, fmin(S_L, 0.0), dL, S_star, F);
    } else if (S_star < 0.0) {
      hllc_star_flux<A>(pr, ig1, irR, S_R, fmax(S_R, 0.0), dR, S_star, F);
    } else {
      double fL[5], fR[5];
      hllc_star_flux<A>(pl, ig1, irL, S_L, fmin(S_L, 0.0), dL, S_star, fL);
      hllc_star_flux<A>(pr, ig1, irR, S_R, fmax(S_R, 0.0), dR, S_star, fR);
#pragma unroll
      for (int v = 0; v < 5; ++v) F[v] = 0.5 * (fL[v] + fR[v]);
    }
#else
    double fL[5] = {0.0, 0.0, 0.0, 0.0, 0.0}, fR[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    if (S_star >= 0.0) hllc_star_flux<A>(pl, ig1, irL, S_L, fmin(S_L, 0.0), dL, S_star, fL);
    if (S_star <= 0.0) hllc_star_flux<A>(pr, ig1, irR, S_R, fmax(S_R, 0.0), dR, S_star, fR);
#pragma unroll
    for (int v = 0; v < 5; ++v) F[v] = (S_star > 0.0) ? fL[v] : ((S_star < 0.0) ? fR[v] : 0.5 * (fL[v] + fR[v]));
#endif
  } else if ((sig >> 11) & 3) {      // HLLC-LM / AUSM+ (out of line, reference order)
    Vec5 a, b;
#pragma unroll
    for (int v = 0; v < 5; ++v) { a.v[v] = pl[v]; b.v[v] = pr[v]; }
    const Vec5 o = riemann_other<A>((sig >> 11) & 3, sig & 15, a, b, gamma);
#pragma unroll
    for (int v = 0; v < 5; ++v) F[v] = o.v[v];
  } else if ((sig >> 4) & 1) {
    // HLL (HLL.py): F = (S_R+ F_L - S_L- F_R + S_L- S_R+ (U_R - U_L)) / (S_R+ - S_L- + eps)
    const int sp = sig & 15;
    const double2 ss = (sp == SIG_EINFELDT) ? einfeldt_signal_speeds(uL, uR, aL, aR, pl[0], pr[0])
                                            : simple_signal_speeds(sp, uL, uR, aL, aR, pl[0], pr[0], pl[4], pr[4], gamma);
    const double wL = fmin(ss.x, 0.0), wR = fmax(ss.y, 0.0);
    double cl[5], cr[5], fl[5], fr[5];
    cons_from_prims_fast(pl, ig1, cl);
    cons_from_prims_fast(pr, ig1, cr);
    physical_flux<A>(pl, cl, fl);
    physical_flux<A>(pr, cr, fr);
    const double inv = rcp_fast(wR - wL + kEps), ww = wL * wR;
#pragma unroll
    for (int v = 0; v < 5; ++v) F[v] = fma(ww, cr[v] - cl[v], fma(wR, fl[v], -(wL * fr[v]))) * inv;
  } else {
    double cl[5], cr[5];
    cons_from_prims_fast(pl, ig1, cl);
    cons_from_prims_fast(pr, ig1, cr);
    const double alpha = fmax(fabs(uL) + aL, fabs(uR) + aR);
    double fl[5], fr[5];
    physical_flux<A>(pl, cl, fl);
    physical_flux<A>(pr, cr, fr);
    const double ha = 0.5 * alpha;
#pragma unroll
    for (int v = 0; v < 5; ++v) F[v] = fma(-ha, cr[v] - cl[v], 0.5 * (fl[v] + fr[v]));
  }
}

// ---------------------------------------------------------------------------
// riemann_flux_main / riemann_flux_exact: the form the software-pipelined sweeps use.  Same formulas as
// riemann_flux, arranged for a SHORT DEPENDENCY CHAIN and NO BRANCH, so that the Riemann solve of face
// j-1 and the reconstruction of face j interleave in one basic block:
//   * a_K = sqrt(gamma p_K) / sqrt(rho_K): the two rsqrt run side by side instead of one after the other;
//   * d_bar = sqrt(N) / (sL + sR) with N = (sL a_L^2 + sR a_R^2)(sL + sR) + sL sR (uR - uL)^2 / 2:
//     rsqrt(N) does not wait for the reciprocal of (sL + sR);
//   * S* = N* / D*, D* = dL - dR < 0 strictly (dL <= -rho_L a_L < 0 < rho_R a_R <= dR), so
//     sign(S*) = -sign(N*) is known BEFORE the division and 1/(S_K - S*) = D* / (S_K D* - N*):
//     the three reciprocals 1/D*, 1/(S_K D* - N*), 1/d_K run side by side;
//   * the side K that sign(S*) selects is chosen by selecting the INPUTS of one star-flux evaluation.
// S* = 0 exactly (sign(0) = 0 in the reference: the flux is the mean of both star fluxes -- every
// face on a symmetry plane) is flagged in `zero`; the caller then replaces F by riemann_flux(), the
// exact branchy form above, in a rarely taken branch at the end of its loop body.
// ---------------------------------------------------------------------------
template <int A, int RIEMANN>
__device__ __forceinline__ void riemann_flux_main(const double (&pl)[5], const double (&pr)[5], double gamma,
                                                  double (&F)[5], bool& zero, int sig = SIG_EINFELDT) {
  using Id = AxisIds<A>;
  if (RIEMANN != RIEMANN_HLLC || (sig & 15) != SIG_EINFELDT) {
    riemann_flux<A, RIEMANN>(pl, pr, gamma, F, sig);
    zero = false;
    return;
  }
  const double ig1 = 1.0 / (gamma - 1.0);
  const double uL = pl[Id::un], uR = pr[Id::un];
  const double gpL = gamma * pl[4], gpR = gamma * pr[4];
  const double yL = rsqrt_fast(pl[0]), yR = rsqrt_fast(pr[0]);       // 1/sqrt(rho)
  const double zL = rsqrt_fast(gpL), zR = rsqrt_fast(gpR);           // 1/sqrt(gamma p)
  const double sL = pl[0] * yL, sR = pr[0] * yR;                     // sqrt(rho)
  const double aL = (gpL * zL) * yL, aR = (gpR * zR) * yR;           // sound speeds
  const double irL = yL * yL, irR = yR * yR;                         // 1/rho
  const double a2L = gpL * irL, a2R = gpR * irR;
  const double ss = sL + sR;
  const double od = rcp_fast(ss);
  const double du = uR - uL;
  const double N = fma(0.5 * (sL * sR), du * du, fma(sL, a2L, sR * a2R) * ss);
  const double d_bar = (N * rsqrt_fast(N)) * od;
  const double u_bar = fma(sL, uL, sR * uR) * od;
  const double S_L = fmin(u_bar - d_bar, uL - aL);
  const double S_R = fmax(u_bar + d_bar, uR + aR);
  const double dL = pl[0] * (S_L - uL);
  const double dR = pr[0] * (S_R - uR);
  const double Ns = (pr[4] - pl[4]) + fma(uL, dL, -(uR * dR));
  const double Ds = dL - dR;                                          // < 0
  const double S_star = Ns * rcp_fast(Ds);
  zero = (Ns == 0.0);
  const bool left = !(Ns > 0.0);                                      // S* >= 0 -> left star state
  double p[5];
#pragma unroll
  for (int v = 0; v < 5; ++v) p[v] = left ? pl[v] : pr[v];
  const double S_K = left ? S_L : S_R;
  const double dK = left ? dL : dR;
  const double irK = left ? irL : irR;
  const double S_lim = left ? fmin(S_L, 0.0) : fmax(S_R, 0.0);
  double c[5];
  cons_from_prims_fast(p, ig1, c);
  const double pre = (dK * Ds) * rcp_fast(fma(S_K, Ds, -Ns));         // rho_K (S_K-u_K)/(S_K-S*)
  const double es = fma(S_star - p[Id::un], fma(p[4], rcp_fast(dK), S_star), c[4] * irK);
  double us[5];
  us[0] = pre;
  us[Id::un] = pre * S_star;
  us[Id::t0] = pre * p[Id::t0];
  us[Id::t1] = pre * p[Id::t1];
  us[4] = pre * es;
  double f[5];
  physical_flux<A>(p, c, f);
#pragma unroll
  for (int v = 0; v < 5; ++v) F[v] = fma(S_lim, us[v] - c[v], f[v]);
}

#endif  // JXF_REFERENCE_ORDER

// ---------------------------------------------------------------------------
// HLLC-LM (HLLCLM.py:30-135: HLLC with the low-Mach wave-speed limiter of Fleischmann et al. 2020, Ma_limit = 0.1)
// and AUSM+ (AUSMP.py:29-95: interface speed of sound ARITHMETIC, alpha = 3/16, beta = 1/8), reference order.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double sign_ref(double x) { return (x > 0.0) ? 1.0 : ((x < 0.0) ? -1.0 : 0.0); }   // jnp.sign

template <int A>
__device__ JXF_NOINLINE Vec5 riemann_other(int variant, int sp, Vec5 PL, Vec5 PR, double gamma) {
  using Id = AxisIds<A>;
  const double (&pl)[5] = PL.v;
  const double (&pr)[5] = PR.v;
  double cl[5], cr[5];
  cons_from_prims(pl, gamma, cl);
  cons_from_prims(pr, gamma, cr);
  const double aL = gsqrt(gdiv(gamma * pl[4], pl[0]));
  const double aR = gsqrt(gdiv(gamma * pr[4], pr[0]));
  const double uL = pl[Id::un], uR = pr[Id::un];
  Vec5 out;
  if (variant == RIEMANN_ALT_HLLCLM) {
    const double2 ss = (sp == SIG_EINFELDT) ? einfeldt_signal_speeds(uL, uR, aL, aR, pl[0], pr[0])
                                            : simple_signal_speeds(sp, uL, uR, aL, aR, pl[0], pr[0], pl[4], pr[4], gamma);
    const double S_L = ss.x, S_R = ss.y;
    const double dL = pl[0] * (S_L - uL);
    const double dR = pr[0] * (S_R - uR);
    const double S_s = gdiv((pr[4] - pl[4]) + (uL * dL - uR * dR), dL - dR);
    double usL[5], usR[5];
    {
      const double pre = gdiv(S_L - uL, S_L - S_s) * pl[0];
      usL[0] = pre;
      usL[Id::un] = pre * S_s;
      usL[Id::t0] = pre * pl[Id::t0];
      usL[Id::t1] = pre * pl[Id::t1];
      usL[4] = pre * (gdiv(cl[4], cl[0]) + (S_s - uL) * (S_s + gdiv(gdiv(pl[4], pl[0]), S_L - uL)));
    }
    {
      const double pre = gdiv(S_R - uR, S_R - S_s) * pr[0];
      usR[0] = pre;
      usR[Id::un] = pre * S_s;
      usR[Id::t0] = pre * pr[Id::t0];
      usR[Id::t1] = pre * pr[Id::t1];
      usR[4] = pre * (gdiv(cr[4], cr[0]) + (S_s - uR) * (S_s + gdiv(gdiv(pr[4], pr[0]), S_R - uR)));
    }
    const double Ma_local = fmax(fabs(gdiv(uL, aL)), fabs(gdiv(uR, aR)));
    const double phi = sin(fmin(1.0, Ma_local / 0.1) * 3.141592653589793 * 0.5);
    const double wL = phi * S_L, wR = phi * S_R;
    double fL[5], fR[5];
    physical_flux<A>(pl, cl, fL);
    physical_flux<A>(pr, cr, fR);
    const double kL = 0.5 * (1.0 + sign_ref(S_L)), kR = 0.5 * (1.0 - sign_ref(S_R));
    const double kS = 0.25 * (1.0 - sign_ref(S_L)) * (1.0 + sign_ref(S_R));
    const double abs_s = fabs(S_s);
#pragma unroll
    for (int v = 0; v < 5; ++v) {
      const double flux_star = 0.5 * (fL[v] + fR[v]) +
                               0.5 * (wL * (usL[v] - cl[v]) + abs_s * (usL[v] - usR[v]) + wR * (usR[v] - cr[v]));
      out.v[v] = kL * fL[v] + kR * fR[v] + kS * flux_star;
    }
  } else {   // RIEMANN_ALT_AUSMP
    const double alpha = 3.0 / 16.0, beta = 1.0 / 8.0;
    const double a = 0.5 * (aL + aR);
    const double M_l = gdiv(uL, a), M_r = gdiv(uR, a);
    const double ql = M_l * M_l - 1.0, qr = M_r * M_r - 1.0;
    const double M_plus = (fabs(M_l) >= 1.0) ? 0.5 * (M_l + fabs(M_l))
                                             : 0.25 * ((M_l + 1.0) * (M_l + 1.0)) + beta * (ql * ql);
    const double M_minus = (fabs(M_r) >= 1.0) ? 0.5 * (M_r - fabs(M_r))
                                              : -0.25 * ((M_r - 1.0) * (M_r - 1.0)) - beta * (qr * qr);
    const double M_ausm = M_plus + M_minus;
    const double M_ausm_plus = 0.5 * (M_ausm + fabs(M_ausm));
    const double M_ausm_minus = 0.5 * (M_ausm - fabs(M_ausm));
    const double P_plus = (fabs(M_l) >= 1.0) ? 0.5 * (1.0 + sign_ref(M_l))
                                             : 0.25 * ((M_l + 1.0) * (M_l + 1.0)) * (2.0 - M_l) + alpha * M_l * (ql * ql);
    const double P_minus = (fabs(M_r) >= 1.0) ? 0.5 * (1.0 - sign_ref(M_r))
                                              : 0.25 * ((M_r - 1.0) * (M_r - 1.0)) * (2.0 + M_r) - alpha * M_r * (qr * qr);
    const double pressure_ausm = P_plus * pl[4] + P_minus * pr[4];
    double phiL[5], phiR[5];
#pragma unroll
    for (int v = 0; v < 4; ++v) { phiL[v] = cl[v]; phiR[v] = cr[v]; }
    phiL[4] = cl[4] + pl[4];
    phiR[4] = cr[4] + pr[4];
#pragma unroll
    for (int v = 0; v < 5; ++v) out.v[v] = a * (M_ausm_plus * phiL[v] + M_ausm_minus * phiR[v]);
    out.v[Id::un] = out.v[Id::un] + pressure_ausm;
  }
  return out;
}

// ---------------------------------------------------------------------------
// Flux-splitting scheme (solvers/convective_fluxes/flux_splitting_scheme.py:62-111) with the conservative
// eigendecomposition of Fedkiw et al. 1999 at the ARITHMETIC frozen state (eigendecomposition.py:146-231, 576-715):
// conservatives U_k and physical fluxes F_k of the six window cells in the characteristic space of the face,
// F+- = (L F_k +- |lambda| L U_k) / 2, F+ reconstructed from the left (j = 0), F- from the right (j = 1), summed and
// transformed back with R.  Eigenvalue magnitudes: ROE :668-671, CLLF :674-681, LLF :684-689.  The conservatives
// of the window cells are formed from the primitives (equation_manager.py:93-101).  Reference order; the matrix
// products run over all five entries in order, zeros included, like the reference's einsum.
// ---------------------------------------------------------------------------
// Right / left eigenvectors of the conservative flux Jacobian at a frozen state (Fedkiw et al. 1999;
// eigendecomposition.py:590-660), as the reference fills them.
template <int A>
__device__ __forceinline__ void conservative_eigenvectors(const Frozen& fz, double (&R)[5][5], double (&L)[5][5]) {
  using Id = AxisIds<A>;
  const double (&ave)[5] = fz.ave;
  const double H = fz.H, G = fz.G, c = fz.c, cc = fz.cc, q2 = fz.q2;
  const double one_cc = grcp(cc), one_rho = grcp(ave[0]);
  const int ua = Id::un, m0 = Id::t0, m1 = Id::t1;
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int k = 0; k < 5; ++k) { R[i][k] = 0.0; L[i][k] = 0.0; }
  R[0][0] = 1.0;
  R[ua][0] = ave[ua] - c;
  R[m0][0] = ave[m0];
  R[m1][0] = ave[m1];
  R[4][0] = H - ave[ua] * c;
  R[0][ua] = G;
  R[1][ua] = G * ave[1];
  R[2][ua] = G * ave[2];
  R[3][ua] = G * ave[3];
  R[4][ua] = G * H - cc;
  R[m0][m0] = -ave[0];
  R[4][m0] = -ave[0] * ave[m0];
  R[m1][m1] = ave[0];
  R[4][m1] = ave[0] * ave[m1];
  R[0][4] = 1.0;
  R[ua][4] = ave[ua] + c;
  R[m0][4] = ave[m0];
  R[m1][4] = ave[m1];
  R[4][4] = H + ave[ua] * c;
  L[0][0] = 0.5 * one_cc * (G * q2 - G * H + (ave[ua] + c) * c);
  L[0][ua] = 0.5 * one_cc * (-ave[ua] * G - c);
  L[0][m0] = 0.5 * one_cc * (-ave[m0] * G);
  L[0][m1] = 0.5 * one_cc * (-ave[m1] * G);
  L[0][4] = 0.5 * one_cc * G;
  L[ua][0] = one_cc * (H - q2);
  L[ua][1] = ave[1] * one_cc;
  L[ua][2] = ave[2] * one_cc;
  L[ua][3] = ave[3] * one_cc;
  L[ua][4] = -one_cc;
  L[m0][0] = ave[m0] * one_rho;
  L[m0][m0] = -one_rho;
  L[m1][0] = -ave[m1] * one_rho;
  L[m1][m1] = one_rho;
  L[4][0] = 0.5 * one_cc * (G * q2 - G * H - (ave[ua] - c) * c);
  L[4][ua] = 0.5 * one_cc * (-ave[ua] * G + c);
  L[4][m0] = 0.5 * one_cc * (-ave[m0] * G);
  L[4][m1] = 0.5 * one_cc * (-ave[m1] * G);
  L[4][4] = 0.5 * one_cc * G;
}

template <int A>
__device__ JXF_NOINLINE Vec5 flux_splitting_flux(Win6 W, double gamma, int id, int fs, int roe) {
  using Id = AxisIds<A>;
  const double (&w)[5][6] = W.w;
  double pL[5], pR[5];
#pragma unroll
  for (int v = 0; v < 5; ++v) {
    pL[v] = w[v][2];
    pR[v] = w[v][3];
  }
  const Frozen fz = frozen_state(pL, pR, gamma, roe);
  const double (&ave)[5] = fz.ave;
  const double c = fz.c;
  const int ua = Id::un;
  double R[5][5], L[5][5];
  conservative_eigenvectors<A>(fz, R, L);
  double lam[5];
  if (fs == FS_ROE) {
    lam[0] = fabs(ave[ua] - c);
    lam[1] = fabs(ave[ua]);
    lam[4] = fabs(ave[ua] + c);
  } else {
    const double cL = gsqrt(gdiv(gamma * pL[4], pL[0])), cR = gsqrt(gdiv(gamma * pR[4], pR[0]));
    if (fs == FS_CLLF) {
      lam[0] = fmax(fabs(pL[ua] - cL), fabs(pR[ua] - cR));
      lam[1] = fmax(fabs(pL[ua]), fabs(pR[ua]));
      lam[4] = fmax(fabs(pL[ua] + cL), fabs(pR[ua] + cR));
    } else {   // FS_LLF
      lam[0] = lam[1] = lam[4] = fmax(fabs(pL[ua]) + cL, fabs(pR[ua]) + cR);
    }
  }
  lam[2] = lam[1];
  lam[3] = lam[1];
  double pos[5][6], neg[5][6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    double p[5], u[5], f[5];
#pragma unroll
    for (int v = 0; v < 5; ++v) p[v] = w[v][k];
    cons_from_prims(p, gamma, u);
    physical_flux<A>(p, u, f);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      double ch = L[i][0] * u[0], cf = L[i][0] * f[0];
#pragma unroll
      for (int v = 1; v < 5; ++v) {
        ch = ch + L[i][v] * u[v];
        cf = cf + L[i][v] * f[v];
      }
      const double lc = lam[i] * ch;
      pos[i][k] = 0.5 * (cf + lc);
      neg[i][k] = 0.5 * (cf - lc);
    }
  }
  double xi[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const double l = stencil_generic(id, 0, pos[i][0], pos[i][1], pos[i][2], pos[i][3], pos[i][4], pos[i][5]);
    const double r = stencil_generic(id, 1, neg[i][5], neg[i][4], neg[i][3], neg[i][2], neg[i][1], neg[i][0]);
    xi[i] = l + r;
  }
  Vec5 out;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    double acc = R[i][0] * xi[0];
#pragma unroll
    for (int v = 1; v < 5; ++v) acc = acc + R[i][v] * xi[v];
    out.v[i] = acc;
  }
  return out;
}

template <int A>
__device__ __forceinline__ void flux_splitting_face(const double (&w)[5][6], double gamma, double (&F)[5], int id, int fs,
                                                    int roe) {
  Win6 W;
#pragma unroll
  for (int v = 0; v < 5; ++v)
#pragma unroll
    for (int k = 0; k < 6; ++k) W.w[v][k] = w[v][k];
  const Vec5 o = flux_splitting_flux<A>(W, gamma, id, fs, roe);
#pragma unroll
  for (int v = 0; v < 5; ++v) F[v] = o.v[v];
}

// reconstruction_variable CONSERVATIVE (high_order_godunov.py:282-296) and CHAR-CONSERVATIVE (:404-417 with
// eigendecomposition.py:576-660, transformtochar / transformtophysical :717-743): the conservatives of the window cells
// (formed from their primitives) are reconstructed as they are, or in the characteristic space of the face's frozen
// state; the face primitives follow from the reconstructed conservatives (equation_manager.py:164-171).
template <int A>
__device__ JXF_NOINLINE Vec10 reconstruct_conservative(Win6 W, double gamma, int id, int mode) {
  const double (&w)[5][6] = W.w;
  double u[5][6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    double p[5], c[5];
#pragma unroll
    for (int v = 0; v < 5; ++v) p[v] = w[v][k];
    cons_from_prims(p, gamma, c);
#pragma unroll
    for (int v = 0; v < 5; ++v) u[v][k] = c[v];
  }
  double cl[5], cr[5];
  if ((mode & 3) == VAR_CONSERVATIVE) {
#pragma unroll
    for (int v = 0; v < 5; ++v) stencil_generic_lr(id, u[v], cl[v], cr[v]);
  } else {
    double pL[5], pR[5];
#pragma unroll
    for (int v = 0; v < 5; ++v) {
      pL[v] = w[v][2];
      pR[v] = w[v][3];
    }
    const Frozen fz = frozen_state(pL, pR, gamma, (mode >> 2) & 1);
    double R[5][5], L[5][5];
    conservative_eigenvectors<A>(fz, R, L);
    double xl[5], xr[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      double ch[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        double acc = L[i][0] * u[0][k];
#pragma unroll
        for (int v = 1; v < 5; ++v) acc = acc + L[i][v] * u[v][k];
        ch[k] = acc;
      }
      stencil_generic_lr(id, ch, xl[i], xr[i]);
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      double al = R[i][0] * xl[0], ar = R[i][0] * xr[0];
#pragma unroll
      for (int v = 1; v < 5; ++v) {
        al = al + R[i][v] * xl[v];
        ar = ar + R[i][v] * xr[v];
      }
      cl[i] = al;
      cr[i] = ar;
    }
  }
  Vec10 o;
  prims_from_cons(cl, gamma, o.l);
  prims_from_cons(cr, gamma, o.r);
  return o;
}

// generic stencil id of the option word: bits 11-14, fifth bit at bit 22
__device__ __forceinline__ int stencil_id(int opt) { return ((opt >> 11) & 15) | (((opt >> 22) & 1) << 4); }

// Interpolation limiter (solvers/positivity/limiter_interpolation.py:77-209, SINGLE-PHASE; eps from
// config/precision.py:54): a reconstructed state whose density is < 1e-12, or whose pressure then is < 1e-10,
// falls back to the first-order state (the adjacent cell: window index 2 for the left, 3 for the right state) --
// density and pressure only (lim = 1) or all primitives (lim = 2, positivity/limit_velocity).  lim = 0: off.
__device__ __forceinline__ void limit_interpolation(double (&p)[5], const double (&w)[5][6], int k, int lim) {
  if (lim == 0) return;
  const bool m1 = p[0] < 1e-12;
  const double p4 = m1 ? w[4][k] : p[4];
  if (m1 || p4 < 1e-10) {
    p[0] = w[0][k];
    p[4] = w[4][k];
    if (lim == 2) {
      p[1] = w[1][k];
      p[2] = w[2][k];
      p[3] = w[3][k];
    }
  }
}

// Positivity-preserving flux limiter (solvers/positivity/limiter_flux.py:146-330, SINGLE-PHASE, flux_limiter
// SIMPLE (mode 1) | NASA (mode 2); eps from config/precision.py:55).  Purely per face: with lambda = dt / dx * sigma
// (sigma = the flux partition of the axis, compute_partition :681-720) the two cells of the face are pseudo-
// integrated, U_minus = U_{i+1} + 2 lambda (F - Fs_{i+1}), U_plus = U_i - 2 lambda (F - Fs_i) (Fs = 0 for SIMPLE,
// the cell's physical flux for NASA); if min density < 1e-12 the face takes the first-order flux (WENO1 states =
// the two cells, HLLC + Einfeldt, :58-99); then the same test on the pressures of the re-integrated states (< 1e-10).
// Out of line and by value: the hot loops only gain a uniform branch on bits 9-10 of the option word.
struct FluxLimArgs {
  const double* dt;     // physical time step size (device scalar; host pointer in the host simulation)
  double inv_dx;        // 1 / dx of the sweep axis
  double sigma;         // flux partition: dim (UNIFORM) or sum_a(1/dx_a) / (1/dx_axis) (CELLSIZE)
};
__device__ __forceinline__ bool below_eps(double a, double b, double eps) {   // jnp.minimum(a, b) < eps (NaN -> false)
  return !(a != a || b != b) && (a < eps || b < eps);
}
template <int A>
__device__ JXF_NOINLINE Vec5 flux_limiter_fix(Vec5 Fin, Vec5 cellL, Vec5 cellR, double gamma, double lam2, int mode) {
  double cL[5], cR[5], Fp[5], F[5], fsL[5] = {0.0, 0.0, 0.0, 0.0, 0.0}, fsR[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int v = 0; v < 5; ++v) F[v] = Fin.v[v];
  cons_from_prims(cellL.v, gamma, cL);
  cons_from_prims(cellR.v, gamma, cR);
  riemann_flux<A, RIEMANN_HLLC>(cellL.v, cellR.v, gamma, Fp, SIG_EINFELDT);
  if (mode == 2) {
    physical_flux<A>(cellL.v, cL, fsL);
    physical_flux<A>(cellR.v, cR, fsR);
  }
  // first integration check: density
  if (below_eps(cR[0] + lam2 * (F[0] - fsR[0]), cL[0] - lam2 * (F[0] - fsL[0]), 1e-12)) {
#pragma unroll
    for (int v = 0; v < 5; ++v) F[v] = Fp[v];
  }
  // second integration check: pressure of the re-integrated states
  double um[5], up[5], wm[5], wp[5];
#pragma unroll
  for (int v = 0; v < 5; ++v) {
    um[v] = cR[v] + lam2 * (F[v] - fsR[v]);
    up[v] = cL[v] - lam2 * (F[v] - fsL[v]);
  }
  prims_from_cons(um, gamma, wm);
  prims_from_cons(up, gamma, wp);
  const bool sw = below_eps(wm[4], wp[4], 1e-10);
  Vec5 out;
#pragma unroll
  for (int v = 0; v < 5; ++v) out.v[v] = sw ? Fp[v] : F[v];
  return out;
}
template <int A>
__device__ __forceinline__ void apply_flux_limiter(const double (&w)[5][6], double gamma, double (&F)[5], int opt,
                                                   const FluxLimArgs& fl) {
  const int mode = (opt >> 9) & 3;
  if (mode == 0) return;
  const double lam2 = 2.0 * (((*fl.dt) * fl.inv_dx) * fl.sigma);
  Vec5 Fin, cl, cr;
#pragma unroll
  for (int v = 0; v < 5; ++v) {
    Fin.v[v] = F[v];
    cl.v[v] = w[v][2];
    cr.v[v] = w[v][3];
  }
  const Vec5 out = flux_limiter_fix<A>(Fin, cl, cr, gamma, lam2, mode);
#pragma unroll
  for (int v = 0; v < 5; ++v) F[v] = out.v[v];
}

// window -> numerical flux at the face (high_order_godunov.py:117-231; flux limiter: space_solver.py:532-543)
template <int A, int RECON, int RIEMANN>
__device__ __forceinline__ void face_flux(const double (&w)[5][6], double gamma, double (&F)[5], int opt,
                                          const FluxLimArgs& fl) {
  if constexpr (RIEMANN == RIEMANN_HLLC_PLAIN) {
    double pl[5], pr[5];
    reconstruct<A, RECON>(w, gamma, pl, pr);
    riemann_flux<A, RIEMANN_HLLC>(pl, pr, gamma, F, SIG_EINFELDT);
    return;
  }
  const int lim = opt & 15, sig = opt >> 4;
  if constexpr ((RECON >> 1) == STENCIL_GENERIC) {
    if ((opt >> 17) & 3) {       // convective_solver = FLUX-SPLITTING
      flux_splitting_face<A>(w, gamma, F, stencil_id(opt), (opt >> 17) & 3, (opt >> 21) & 1);
      return;
    }
  }
  double pl[5], pr[5];
  reconstruct<A, RECON>(w, gamma, pl, pr, stencil_id(opt), (opt >> 19) & 7);
  limit_interpolation(pl, w, 2, lim);
  limit_interpolation(pr, w, 3, lim);
  riemann_flux<A, RIEMANN>(pl, pr, gamma, F, sig);
  apply_flux_limiter<A>(w, gamma, F, opt, fl);
}

#ifndef JXF_REFERENCE_ORDER
// the option-free face flux with the as-is fields' weights given (sweep_rows' lane carry); tuned stencils, HLLC_PLAIN
template <int A, int RECON>
__device__ __forceinline__ void face_flux_given(const double (&w)[5][6], double gamma, double (&F)[5],
                                                const ReconCarry<RECON>& gl, const ReconCarry<RECON>& gr) {
  double pl[5], pr[5];
  reconstruct_given<A, RECON>(w, gamma, pl, pr, gl, gr);
  riemann_flux<A, RIEMANN_HLLC>(pl, pr, gamma, F, SIG_EINFELDT);
}
#endif

#ifdef JXF_REFERENCE_ORDER
template <int A, int RIEMANN>
__device__ __forceinline__ void riemann_flux_main(const double (&pl)[5], const double (&pr)[5], double gamma,
                                                  double (&F)[5], bool& zero, int sig = SIG_EINFELDT) {
  riemann_flux<A, RIEMANN>(pl, pr, gamma, F, sig);
  zero = false;
}
template <int RECON>
struct ReconCarry {};
template <int A, int RECON>
__device__ __forceinline__ void recon_carry_init(const double (&)[5][6], ReconCarry<RECON>&) {}
template <int A, int RECON>
__device__ __forceinline__ void reconstruct_carry(const double (&w)[5][6], double gamma, double (&pl)[5],
                                                  double (&pr)[5], ReconCarry<RECON>&, int alt = 0, int mode = 0) {
  reconstruct<A, RECON>(w, gamma, pl, pr, alt, mode);
}
template <int A, int RECON, int RIEMANN>
__device__ __forceinline__ void face_flux_carry(const double (&w)[5][6], double gamma, double (&F)[5], ReconCarry<RECON>&,
                                                int opt, const FluxLimArgs& fl) {
  face_flux<A, RECON, RIEMANN>(w, gamma, F, opt, fl);
}
#else
// marching variant: shares the cell-centred weights of the as-is fields between consecutive faces
template <int A, int RECON, int RIEMANN>
__device__ __forceinline__ void face_flux_carry(const double (&w)[5][6], double gamma, double (&F)[5],
                                                ReconCarry<RECON>& cy, int opt, const FluxLimArgs& fl) {
  if constexpr (RIEMANN == RIEMANN_HLLC_PLAIN) {
    double pl[5], pr[5];
    reconstruct_carry<A, RECON>(w, gamma, pl, pr, cy);
    riemann_flux<A, RIEMANN_HLLC>(pl, pr, gamma, F, SIG_EINFELDT);
    return;
  }
  const int lim = opt & 15, sig = opt >> 4;
  if constexpr ((RECON >> 1) == STENCIL_GENERIC) {
    if ((opt >> 17) & 3) {       // convective_solver = FLUX-SPLITTING
      flux_splitting_face<A>(w, gamma, F, stencil_id(opt), (opt >> 17) & 3, (opt >> 21) & 1);
      return;
    }
  }
  double pl[5], pr[5];
  reconstruct_carry<A, RECON>(w, gamma, pl, pr, cy, stencil_id(opt), (opt >> 19) & 7);
  limit_interpolation(pl, w, 2, lim);
  limit_interpolation(pr, w, 3, lim);
#if JXF_RIEMANN_MAIN
  bool zero;
  riemann_flux_main<A, RIEMANN>(pl, pr, gamma, F, zero, sig);
  if (zero) riemann_flux<A, RIEMANN>(pl, pr, gamma, F, sig);
#else
  riemann_flux<A, RIEMANN>(pl, pr, gamma, F, sig);
#endif
  apply_flux_limiter<A>(w, gamma, F, opt, fl);
}
#endif

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
#ifndef JXF_REFERENCE_ORDER
  // the same with the MUFU + Newton reciprocal / rsqrt (<= 3 ulp on c; dt = CFL dx / (max + eps) moves by as much) and
  // plain compare-selects (finite data); used by the tuned epilogue instantiations
  __device__ __forceinline__ void add_cell_fast(const double (&p)[5], double gamma, int active_mask) {
    const double x = gamma * p[4] * rcp_fast(p[0]);
    const double c = x * rsqrt_fast(x);
    double s = 0.0;
    if (active_mask & 1) s += fabs(p[1]) + c;
    if (active_mask & 2) s += fabs(p[2]) + c;
    if (active_mask & 4) s += fabs(p[3]) + c;
    max_s = (s > max_s) ? s : max_s;
    min_rho = (p[0] < min_rho) ? p[0] : min_rho;
    min_p = (p[4] < min_p) ? p[4] : min_p;
  }
#endif
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
