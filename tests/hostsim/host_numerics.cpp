// Host build of jaxfluids_b200/csrc/numerics.cuh (TEST TOOL, not product): lets the CPU-only
// container check the arithmetic of the device functions (with FMA contraction, like nvcc's
// default -fmad=true) against the oracle before spending GPU time.
#define __device__
#define __host__
#define __forceinline__ inline
#include <cmath>
using std::fabs; using std::fmin; using std::fmax; using std::sqrt; using std::fma;
#include "../../jaxfluids_b200/csrc/numerics.cuh"

using namespace jxf;

template <int A, int RECON, int RIEMANN>
static void run(const double* win, long n, double gamma, double* out, int opt, const FluxLimArgs& fl) {
  for (long i = 0; i < n; ++i) {
    double w[5][6], F[5];
    for (int v = 0; v < 5; ++v) for (int k = 0; k < 6; ++k) w[v][k] = win[(i * 5 + v) * 6 + k];
    face_flux<A, RECON, RIEMANN>(w, gamma, F, opt, fl);
    for (int v = 0; v < 5; ++v) out[i * 5 + v] = F[v];
  }
}

// marching variant: `nseq` sequences of `len` consecutive faces each (windows laid out (nseq, len, 5, 6));
// the carry is primed from the first window of each sequence, exactly as sweep_strided does
template <int A, int RECON, int RIEMANN>
static void run_march(const double* win, long nseq, long len, double gamma, double* out, int opt,
                      const FluxLimArgs& fl) {
  for (long s = 0; s < nseq; ++s) {
    ReconCarry<RECON> cy;
    for (long i = 0; i < len; ++i) {
      double w[5][6], F[5];
      const double* p = win + ((s * len + i) * 5) * 6;
      for (int v = 0; v < 5; ++v) for (int k = 0; k < 6; ++k) w[v][k] = p[v * 6 + k];
      if (i == 0) recon_carry_init<A, RECON>(w, cy);
      face_flux_carry<A, RECON, RIEMANN>(w, gamma, F, cy, opt, fl);
      for (int v = 0; v < 5; ++v) out[(s * len + i) * 5 + v] = F[v];
    }
  }
}

extern "C" int face_flux_march_host(int axis, int recon, int riemann, const double* win, long nseq, long len, double gamma,
                                    double* out, int opt, double dt, double inv_dx, double sigma) {
  const FluxLimArgs fl = {&dt, inv_dx, sigma};
#define MCASE(A, R, S) if (axis == A && recon == R && riemann == S) { run_march<A, R, S>(win, nseq, len, gamma, out, opt, fl); return 0; }
  MCASE(0,0,0) MCASE(0,0,1) MCASE(0,1,0) MCASE(0,1,1) MCASE(0,2,0) MCASE(0,2,1) MCASE(0,3,0) MCASE(0,3,1)
  MCASE(1,0,0) MCASE(1,0,1) MCASE(1,1,0) MCASE(1,1,1) MCASE(1,2,0) MCASE(1,2,1) MCASE(1,3,0) MCASE(1,3,1)
  MCASE(2,0,0) MCASE(2,0,1) MCASE(2,1,0) MCASE(2,1,1) MCASE(2,2,0) MCASE(2,2,1) MCASE(2,3,0) MCASE(2,3,1)
  MCASE(0,4,0) MCASE(0,4,1) MCASE(0,5,0) MCASE(0,5,1) MCASE(1,4,0) MCASE(1,4,1) MCASE(1,5,0) MCASE(1,5,1)
  MCASE(2,4,0) MCASE(2,4,1) MCASE(2,5,0) MCASE(2,5,1)
  return -1;
}

// windows: (n, 5, 6) doubles; out: (n, 5)
extern "C" int face_flux_host(int axis, int recon, int riemann, const double* win, long n, double gamma, double* out, int opt,
                              double dt, double inv_dx, double sigma) {
  const FluxLimArgs fl = {&dt, inv_dx, sigma};
#define CASE(A, R, S) if (axis == A && recon == R && riemann == S) { run<A, R, S>(win, n, gamma, out, opt, fl); return 0; }
  CASE(0,0,0) CASE(0,0,1) CASE(0,1,0) CASE(0,1,1) CASE(0,2,0) CASE(0,2,1) CASE(0,3,0) CASE(0,3,1)
  CASE(1,0,0) CASE(1,0,1) CASE(1,1,0) CASE(1,1,1) CASE(1,2,0) CASE(1,2,1) CASE(1,3,0) CASE(1,3,1)
  CASE(2,0,0) CASE(2,0,1) CASE(2,1,0) CASE(2,1,1) CASE(2,2,0) CASE(2,2,1) CASE(2,3,0) CASE(2,3,1)
  CASE(0,4,0) CASE(0,4,1) CASE(0,5,0) CASE(0,5,1) CASE(1,4,0) CASE(1,4,1) CASE(1,5,0) CASE(1,5,1)
  CASE(2,4,0) CASE(2,4,1) CASE(2,5,0) CASE(2,5,1)
  return -1;
}
