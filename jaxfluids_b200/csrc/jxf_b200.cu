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
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <cmath>
#include <new>

#include "../../include/jxf_b200.h"
#include "numerics.cuh"

namespace jxf {

struct Geom {
  int n[3];            // interior cells
  int ext[3];          // buffer extents (n + 2nh, or 1)
  int off[3];          // nh on active axes, 0 on inactive
  int nh;
  long long st[3];     // element strides of the halo'd buffers
  long long vst;       // variable stride of the halo'd buffers
  long long rst[3];    // element strides of the interior-only rhs buffer
  long long rvst;
};

struct SweepArgs {
  const double* prims;     // stage-entry primitives (halo'd)
  double* rhs;             // interior-only accumulator
  const double* cons_in;   // EPI only
  const double* cons_n;    // EPI only, stage > 0
  double* cons_out;        // EPI only
  double* prims_out;       // EPI only
  const double* dt;        // EPI only (device scalar)
  double* red;             // EPI only, 3 doubles
  double ca, cb;           // RK blend U = ca*U + cb*U^n
  double dt_mult;          // RK stage dt multiplier
  double gamma;
  double inv_dx;
  int blend;               // stage > 0
  int has_prev;            // EPI: rhs holds earlier axes' sum
  int accumulate;          // !EPI: rhs += (1) or rhs = 0.0 + (0)
  int reduce;              // EPI: update red
  int active_mask;         // bit i = axis i active
  int chunk_len;           // strided: cells per chunk along A
  int span;                // contig: faces per range
  int range_lo, range_hi;  // strided: cells [range_lo, range_hi) along A are swept (default: all)
  int fuse_halo;           // EPI: also write the outer-BC halo images of boundary-adjacent cells
  int nh;
  int bc[6];               // JXF_BC_* per physical face (east,west,north,south,top,bottom)
  int limiter;             // face-flux options: interpolation limiter (0 off, 1 density + pressure, 2 all primitives)
                           // | signal speed (JXF_SIGNAL_*) << 4 | HLL solver << 8 | flux limiter (1 SIMPLE, 2 NASA) << 9
  FluxLimArgs fl;          // positivity flux limiter: dt pointer, 1/dx of the axis, flux partition
  int volume_force;        // EPI: add the gravity source (g_i rho, g . rho u) of the stage's conservatives
  double gravity[3];
  double wall[6][3];       // wall velocity (u, v, w) per JXF_BC_WALL face
  double dirichlet[6][5];  // prescribed primitives per JXF_BC_DIRICHLET face
};

// ---------------------------------------------------------------------------
// sweep geometry in ROLES (filled on the host): A = sweep axis; T1/T2 = the two transverse axes
// with T2 the faster one in memory.  Strided sweeps: T2 is the contiguous axis (lanes run along
// it).  Contiguous sweeps: A itself is the contiguous axis and rows are indexed (i1, i2).
// All offsets are relative to the first INTERIOR cell (h0 is folded into the base pointers).
// ---------------------------------------------------------------------------
struct SweepGeom {
  int axA, ax1, ax2;         // physical axis of each role
  int bcA_hi, bcA_lo, bc1_hi, bc1_lo, bc2_hi, bc2_lo;   // JXF_BC_* of the faces of each role (halo fusion)
  int nA, n1, n2;
  long long sA, s1, s2;      // strides in the halo'd buffers
  long long rA, r1, r2;      // strides in the interior-only rhs buffer
  long long vst, rvst;       // variable strides
};

#ifndef JXF_MIN_BLOCKS
#define JXF_MIN_BLOCKS 3
#endif
#ifndef JXF_PREFETCH
#define JXF_PREFETCH 0
#endif
#ifndef JXF_ROWS_KERNEL
#define JXF_ROWS_KERNEL 1
#endif

__device__ __forceinline__ void prefetch_l2(const double* p) {
#if JXF_PREFETCH == 2
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}

// operands of the cell update that come from memory; loaded EARLY (before the flux arithmetic of
// the iteration) so their latency hides behind ~700 FP64 instructions
template <int EPI>
struct CellIn {
  double rhs[5];   // EPI=0: accumulate target (if accumulate); EPI=1: earlier axes' sum (if has_prev)
  double U[5];     // EPI=1
  double Un[5];    // EPI=1, blend
};

template <int EPI>
__device__ __forceinline__ void load_cell_in(const SweepGeom& g, const SweepArgs& a, long long hidx, long long ridx,
                                             CellIn<EPI>& in) {
  if (EPI == 0) {
    if (a.accumulate) {
#pragma unroll
      for (int v = 0; v < 5; ++v) in.rhs[v] = a.rhs[ridx + v * g.rvst];
    }
  } else {
#pragma unroll
    for (int v = 0; v < 5; ++v) in.U[v] = a.cons_in[hidx + v * g.vst];
    if (a.has_prev) {
#pragma unroll
      for (int v = 0; v < 5; ++v) in.rhs[v] = a.rhs[ridx + v * g.rvst];
    }
    if (a.blend) {
#pragma unroll
      for (int v = 0; v < 5; ++v) in.Un[v] = a.cons_n[hidx + v * g.vst];
    }
  }
}

// Outer-BC halo images of one freshly updated interior cell (halos/outer/material.py:868-894,
// boundary_condition.py:563-595, :698-731), fused into the stage epilogue: every face-halo cell of
// PERIODIC / SYMMETRY / ZEROGRADIENT faces is the image of exactly one interior cell within nh of
// that face, so the thread that produced the cell also writes its images (prims, and cons
// recomputed from the image prims, :248-250).  i = interior index along the role axis.
struct HaloOut {
  double* prims;
  double* cons;
  long long vst;
  double gamma;
  int nh;
};

// wall = nullptr: copy, negating velocity component flip_var (1..3; -1 = none).  wall != nullptr: no-slip wall
// moving with (u, v, w) = wall[0..2]: every velocity component becomes 2 u_wall - u (halos/outer/material.py:510-512)
__device__ __forceinline__ void halo_image(const HaloOut& h, long long dst, double p0, double p1, double p2, double p3,
                                           double p4, int flip_var, const double* wall = nullptr) {
  double q[5] = {p0, p1, p2, p3, p4};
  if (wall) {
#pragma unroll
    for (int v = 1; v < 4; ++v) q[v] = 2 * wall[v - 1] - q[v];
  } else {
#pragma unroll
    for (int v = 1; v < 4; ++v) q[v] = (v == flip_var) ? q[v] * -1.0 : q[v];
  }
  double c[5];
  cons_from_prims(q, h.gamma, c);
#pragma unroll
  for (int v = 0; v < 5; ++v) {
    h.prims[dst + v * h.vst] = q[v];
    h.cons[dst + v * h.vst] = c[v];
  }
}

// one role axis of the images of a cell; wall_hi / wall_lo: wall velocities of the two faces of this axis
__device__ __forceinline__ void halo_images_axis(const HaloOut& h, int bhi, int blo, long long hidx, const double (&p)[5],
                                                 int ax, int n, int i, long long stride, const double* wall_hi,
                                                 const double* wall_lo, const double* dir_hi, const double* dir_lo) {
  if (n <= 1) return;
  const int nh = h.nh;
  // low side (west / south / bottom)
  if (blo == JXF_BC_SYMMETRY) {
    if (i < nh) halo_image(h, hidx + (long long)(-1 - 2 * i) * stride, p[0], p[1], p[2], p[3], p[4], 1 + ax);
  } else if (blo == JXF_BC_WALL) {
    if (i < nh) halo_image(h, hidx + (long long)(-1 - 2 * i) * stride, p[0], p[1], p[2], p[3], p[4], -1, wall_lo);
  } else if (blo == JXF_BC_PERIODIC) {
    if (i >= n - nh) halo_image(h, hidx - (long long)n * stride, p[0], p[1], p[2], p[3], p[4], -1);
  } else if (blo == JXF_BC_ZEROGRADIENT) {
    if (i == 0)
      for (int l = 1; l <= nh; ++l) halo_image(h, hidx - (long long)l * stride, p[0], p[1], p[2], p[3], p[4], -1);
  } else if (blo == JXF_BC_DIRICHLET) {       // constants: written by the thread of the boundary-adjacent cell
    if (i == 0)
      for (int l = 1; l <= nh; ++l)
        halo_image(h, hidx - (long long)l * stride, dir_lo[0], dir_lo[1], dir_lo[2], dir_lo[3], dir_lo[4], -1);
  }
  // high side (east / north / top)
  if (bhi == JXF_BC_SYMMETRY) {
    if (i >= n - nh) halo_image(h, hidx + (long long)(2 * (n - i) - 1) * stride, p[0], p[1], p[2], p[3], p[4], 1 + ax);
  } else if (bhi == JXF_BC_WALL) {
    if (i >= n - nh) halo_image(h, hidx + (long long)(2 * (n - i) - 1) * stride, p[0], p[1], p[2], p[3], p[4], -1, wall_hi);
  } else if (bhi == JXF_BC_PERIODIC) {
    if (i < nh) halo_image(h, hidx + (long long)n * stride, p[0], p[1], p[2], p[3], p[4], -1);
  } else if (bhi == JXF_BC_ZEROGRADIENT) {
    if (i == n - 1)
      for (int l = 1; l <= nh; ++l) halo_image(h, hidx + (long long)l * stride, p[0], p[1], p[2], p[3], p[4], -1);
  } else if (bhi == JXF_BC_DIRICHLET) {
    if (i == n - 1)
      for (int l = 1; l <= nh; ++l)
        halo_image(h, hidx + (long long)l * stride, dir_hi[0], dir_hi[1], dir_hi[2], dir_hi[3], dir_hi[4], -1);
  }
}

// OUT OF LINE on purpose: executed only by the thin shell of boundary-adjacent cells; keeping it out of
// the sweep loop keeps the hot loop short (instruction cache) and its register allocation free of this code
__device__ __noinline__ void halo_images_cell(const SweepGeom& g, const SweepArgs& a, long long hidx, double p0, double p1,
                                              double p2, double p3, double p4, int iA, int i1, int i2);

template <int EPI>
__device__ __forceinline__ void finalize_cell(const SweepGeom& g, const SweepArgs& a, long long hidx, long long ridx,
                                              const CellIn<EPI>& in, const double (&r)[5], double step, Red& red,
                                              int iA, int i1, int i2) {
  // r = F_{i-1/2} - F_{i+1/2}; the axis contribution (1/dx) r (space_solver.py:597-599) is added to the
  // earlier axes' sum with one fused multiply-add
  if (EPI == 0) {
#pragma unroll
    for (int v = 0; v < 5; ++v) a.rhs[ridx + v * g.rvst] = a.accumulate ? fma(a.inv_dx, r[v], in.rhs[v]) : a.inv_dx * r[v];
  } else {
    double U[5];
#pragma unroll
    for (int v = 0; v < 5; ++v) {
      double tot = a.has_prev ? fma(a.inv_dx, r[v], in.rhs[v]) : a.inv_dx * r[v];
      if (a.volume_force) {     // space_solver.py:378-384 with the einsums of source_term_solver.py:180-182
        if (v >= 1 && v <= 3) tot += a.gravity[v - 1] * in.U[0];
        if (v == 4) tot += (a.gravity[0] * in.U[1] + a.gravity[1] * in.U[2]) + a.gravity[2] * in.U[3];
      }
      double u = in.U[v];
      if (a.blend) u = a.ca * u + a.cb * in.Un[v];
      U[v] = u + step * tot;
    }
    double p[5];
    prims_from_cons(U, a.gamma, p);
#pragma unroll
    for (int v = 0; v < 5; ++v) {
      a.cons_out[hidx + v * g.vst] = U[v];
      a.prims_out[hidx + v * g.vst] = p[v];
    }
    if (a.reduce) red.add_cell(p, a.gamma, a.active_mask);
    if (a.fuse_halo) {
      // boundary-adjacent cells only (a thin shell); warp-divergent by construction
      const bool near = (iA < a.nh) | (iA >= g.nA - a.nh) | (i1 < a.nh) | (i1 >= g.n1 - a.nh) | (i2 < a.nh) |
                        (i2 >= g.n2 - a.nh);
      if (near) halo_images_cell(g, a, hidx, p[0], p[1], p[2], p[3], p[4], iA, i1, i2);
    }
  }
}

__device__ __noinline__ void halo_images_cell(const SweepGeom& g, const SweepArgs& a, long long hidx, double p0, double p1,
                                              double p2, double p3, double p4, int iA, int i1, int i2) {
  const HaloOut h{a.prims_out, a.cons_out, g.vst, a.gamma, a.nh};
  const double p[5] = {p0, p1, p2, p3, p4};
  halo_images_axis(h, g.bcA_hi, g.bcA_lo, hidx, p, g.axA, g.nA, iA, g.sA, a.wall[2 * g.axA], a.wall[2 * g.axA + 1],
                   a.dirichlet[2 * g.axA], a.dirichlet[2 * g.axA + 1]);
  halo_images_axis(h, g.bc1_hi, g.bc1_lo, hidx, p, g.ax1, g.n1, i1, g.s1, a.wall[2 * g.ax1], a.wall[2 * g.ax1 + 1],
                   a.dirichlet[2 * g.ax1], a.dirichlet[2 * g.ax1 + 1]);
  halo_images_axis(h, g.bc2_hi, g.bc2_lo, hidx, p, g.ax2, g.n2, i2, g.s2, a.wall[2 * g.ax2], a.wall[2 * g.ax2 + 1],
                   a.dirichlet[2 * g.ax2], a.dirichlet[2 * g.ax2 + 1]);
}

#ifdef JXF_WITH_STRIDED   // the register-window predecessor of sweep_march: A/B builds only (-DJXF_WITH_STRIDED)
// ---------------------------------------------------------------------------
// strided sweep: thread = one (i1, i2) column (i2 along the contiguous axis), marching along A
// over one chunk with a rolling 6-cell register window; each face flux is computed once.
// ---------------------------------------------------------------------------
template <int A, int RECON, int RIEMANN, int EPI>
__global__ void __launch_bounds__(128, JXF_MIN_BLOCKS) sweep_strided(const __grid_constant__ SweepGeom g, const __grid_constant__ SweepArgs a) {
  const long long plane = (long long)g.n1 * g.n2;
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  Red red;
  red.init();
  if (p < plane) {
    const int i1 = (int)(p / g.n2);
    const int i2 = (int)(p - (long long)i1 * g.n2);
    const int f0 = a.range_lo + blockIdx.y * a.chunk_len;
    const int f1 = min(f0 + a.chunk_len, a.range_hi);
    const long long sA = g.sA;
    const long long col_h = i1 * g.s1 + i2 * g.s2;
    const long long col_r = i1 * g.r1 + i2 * g.r2;
    const double* base = a.prims + col_h + (long long)(f0 - 3) * sA;   // cell f0-3
    const double step = (EPI ? (*a.dt) * a.dt_mult : 0.0);
    double w[5][6], nx[5], Fp[5];
#pragma unroll
    for (int v = 0; v < 5; ++v) {
#pragma unroll
      for (int k = 0; k < 5; ++k) w[v][k] = base[v * g.vst + k * sA];
      nx[v] = base[v * g.vst + 5 * sA];
      Fp[v] = 0.0;
    }
    ReconCarry<RECON> cy;
    recon_carry_init<A, RECON>(w, cy);        // weights of cell f0-1 (left stencil of the first face)
    for (int f = f0; f <= f1; ++f) {
#pragma unroll
      for (int v = 0; v < 5; ++v) w[v][5] = nx[v];
      if (f < f1) {   // prefetch cell f+3 (<= n+2 < n+nh since nh >= 3) for the next iteration
        const double* nb = base + (long long)(f - f0 + 6) * sA;
#pragma unroll
        for (int v = 0; v < 5; ++v) nx[v] = nb[v * g.vst];
      }
      const long long hidx = col_h + (long long)(f - 1) * sA;
      const long long ridx = col_r + (long long)(f - 1) * g.rA;
      CellIn<EPI> in;
      if (f > f0) load_cell_in<EPI>(g, a, hidx, ridx, in);
      double F[5];
      face_flux_carry<A, RECON, RIEMANN>(w, a.gamma, F, cy, a.limiter, a.fl);
      if (f > f0) {
        double r[5];
#pragma unroll
        for (int v = 0; v < 5; ++v) r[v] = Fp[v] - F[v];
        finalize_cell<EPI>(g, a, hidx, ridx, in, r, step, red, f - 1, i1, i2);
      }
#pragma unroll
      for (int v = 0; v < 5; ++v) {
        Fp[v] = F[v];
#pragma unroll
        for (int k = 0; k < 5; ++k) w[v][k] = w[v][k + 1];
      }
    }
  }
  if (EPI) {
    if (a.reduce) red_commit(red, a.red);
  }
}

#endif  // JXF_WITH_STRIDED

// ---------------------------------------------------------------------------
// strided sweep, production form ("march"): same thread mapping as sweep_strided, but the 6-cell
// window lives in a per-thread column of a shared-memory RING of planes instead of registers.
// Every iteration each thread posts ONE asynchronous copy (cp.async, 8 B x 5 variables, coalesced
// across the warp) of the plane kRingAhead steps ahead of the window straight from global to shared
// memory -- no staging registers, no window shift (the ring index rotates instead of the data) --
// and reads the window values where the arithmetic needs them.  A thread only ever touches its own
// column of the ring, so the pipeline needs no barrier: cp.async.wait_group orders a thread's own
// copies.  Freed registers (~70 of 168) go to instruction-level parallelism of the FP64 arithmetic.
// ---------------------------------------------------------------------------
#ifndef JXF_MARCH_BLOCKS
#define JXF_MARCH_BLOCKS 3
#endif
#ifndef JXF_MARCH_KERNEL
#define JXF_MARCH_KERNEL 1
#endif
constexpr int kRingSlots = 8;                   // planes in the ring (power of two)
constexpr int kRingAhead = kRingSlots - 6;      // planes in flight beyond the current window

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ring_copy8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void ring_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void ring_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int A, int RECON, int RIEMANN, int EPI>
__global__ void __launch_bounds__(128, JXF_MARCH_BLOCKS)
sweep_march(const __grid_constant__ SweepGeom g, const __grid_constant__ SweepArgs a) {
  __shared__ double ring[kRingSlots][5][128];
  const int t = threadIdx.x;
  const long long plane = (long long)g.n1 * g.n2;
  const long long p = blockIdx.x * (long long)blockDim.x + t;
  Red red;
  red.init();
  if (p < plane) {
    const int i1 = (int)(p / g.n2);
    const int i2 = (int)(p - (long long)i1 * g.n2);
    const int f0 = a.range_lo + blockIdx.y * a.chunk_len;
    const int f1 = min(f0 + a.chunk_len, a.range_hi);
    const long long sA = g.sA;
    const long long col_h = i1 * g.s1 + i2 * g.s2;
    const long long col_r = i1 * g.r1 + i2 * g.r2;
    const double* base = a.prims + col_h + (long long)(f0 - 3) * sA;   // ring cell 0 = cell f0-3
    const int last_cell = f1 - f0 + 5;                                  // ring cell of cell f1+2 (< n+nh: nh >= 3)
    const double step = (EPI ? (*a.dt) * a.dt_mult : 0.0);
    auto post = [&](int c) {
      if (c <= last_cell) {
        const double* src = base + (long long)c * sA;
#pragma unroll
        for (int v = 0; v < 5; ++v) ring_copy8(&ring[c & (kRingSlots - 1)][v][t], src + v * g.vst);
      }
      ring_commit();
    };
#pragma unroll
    for (int c = 0; c < 5 + kRingAhead; ++c) post(c);
    ring_wait<kRingAhead>();                   // cells 0..4 have landed
    ReconCarry<RECON> cy;
    {
      double w[5][6];
#pragma unroll
      for (int v = 0; v < 5; ++v) {
#pragma unroll
        for (int k = 0; k < 5; ++k) w[v][k] = ring[k][v][t];
        w[v][5] = 0.0;
      }
      recon_carry_init<A, RECON>(w, cy);       // weights of cell f0-1 (left stencil of the first face)
    }
    double Fp[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    const int nfaces = f1 - f0 + 1;
    for (int j = 0; j < nfaces; ++j) {
      post(j + 5 + kRingAhead);                // overwrites the slot of cell j-1, which no window needs any more
      const int f = f0 + j;
      const long long hidx = col_h + (long long)(f - 1) * sA;
      const long long ridx = col_r + (long long)(f - 1) * g.rA;
      CellIn<EPI> in;
      if (j > 0) load_cell_in<EPI>(g, a, hidx, ridx, in);
      ring_wait<kRingAhead>();                 // cells j..j+5 have landed
      double w[5][6];
#pragma unroll
      for (int v = 0; v < 5; ++v)
#pragma unroll
        for (int k = 0; k < 6; ++k) w[v][k] = ring[(j + k) & (kRingSlots - 1)][v][t];
      double F[5];
      face_flux_carry<A, RECON, RIEMANN>(w, a.gamma, F, cy, a.limiter, a.fl);
      if (j > 0) {
        double r[5];
#pragma unroll
        for (int v = 0; v < 5; ++v) r[v] = Fp[v] - F[v];
        finalize_cell<EPI>(g, a, hidx, ridx, in, r, step, red, f - 1, i1, i2);
      }
#pragma unroll
      for (int v = 0; v < 5; ++v) Fp[v] = F[v];
    }
    ring_wait<0>();
  }
  if (EPI) {
    if (a.reduce) red_commit(red, a.red);
  }
}

// ---------------------------------------------------------------------------
// contiguous sweep: lanes = consecutive faces of the flattened (row, face) sequence; the left
// face flux comes from lane-1 by shuffle (lane 0: carry from the warp's previous iteration).
// ---------------------------------------------------------------------------
template <int A, int RECON, int RIEMANN, int EPI>
__global__ void __launch_bounds__(128, JXF_MIN_BLOCKS) sweep_contig(const __grid_constant__ SweepGeom g, const __grid_constant__ SweepArgs a,
                                                                    const long long total_faces) {
  const int nf = g.nA + 1;
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long nranges = (total_faces + a.span - 1) / a.span;
  const double step = (EPI ? (*a.dt) * a.dt_mult : 0.0);
  Red red;
  red.init();
  for (long long range = warp; range < nranges; range += nwarps) {
    const long long gs = range * a.span;
    const long long ge = min(gs + (long long)a.span, total_faces);
    // one carry-in face unless the range starts a row
    const long long gbeg = (gs % nf == 0) ? gs : gs - 1;
    const int iters = (int)((ge - gbeg + 31) >> 5);
    double carry[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    long long gf = gbeg + lane;
    long long row = gf / nf;
    int f = (int)(gf - row * nf);
    for (int it = 0; it < iters; ++it) {
      const bool act = gf < ge;
      const bool fin = act && f > 0 && gf > gbeg;
      double F[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
      long long hidx = 0, ridx = 0;
      int i1 = 0, i2 = 0;
      CellIn<EPI> in;
      if (act) {
        i1 = (int)(row / g.n2);
        i2 = (int)(row - (long long)i1 * g.n2);
        const long long col_h = i1 * g.s1 + i2 * g.s2;
        hidx = col_h + (long long)(f - 1) * g.sA;
        ridx = i1 * g.r1 + i2 * g.r2 + (long long)(f - 1) * g.rA;
        const double* base = a.prims + col_h + (long long)(f - 3) * g.sA;
#if JXF_PREFETCH
        if (it + 1 < iters) {
          // pull the next iteration's lines (32 faces further along the flattened row sequence;
          // rows are contiguous in memory up to the halo gap) towards the SM while this one computes
          const long long nxt = 32 * g.sA;
#pragma unroll
          for (int v = 0; v < 5; ++v) prefetch_l2(base + v * g.vst + nxt + 3 * g.sA);
          if (EPI) {
#pragma unroll
            for (int v = 0; v < 5; ++v) prefetch_l2(a.cons_in + hidx + v * g.vst + nxt);
            if (a.has_prev) {
#pragma unroll
              for (int v = 0; v < 5; ++v) prefetch_l2(a.rhs + ridx + v * g.rvst + 32 * g.rA);
            }
            if (a.blend) {
#pragma unroll
              for (int v = 0; v < 5; ++v) prefetch_l2(a.cons_n + hidx + v * g.vst + nxt);
            }
          } else if (a.accumulate) {
#pragma unroll
            for (int v = 0; v < 5; ++v) prefetch_l2(a.rhs + ridx + v * g.rvst + 32 * g.rA);
          }
        }
#endif
        double w[5][6];
#pragma unroll
        for (int v = 0; v < 5; ++v)
#pragma unroll
          for (int k = 0; k < 6; ++k) w[v][k] = base[v * g.vst + k * g.sA];
        if (fin) load_cell_in<EPI>(g, a, hidx, ridx, in);
        face_flux<A, RECON, RIEMANN>(w, a.gamma, F, a.limiter, a.fl);
      }
      double Fl[5];
#pragma unroll
      for (int v = 0; v < 5; ++v) {
        const double up = __shfl_up_sync(0xffffffffu, F[v], 1);
        const double last = __shfl_sync(0xffffffffu, F[v], 31);
        Fl[v] = (lane == 0) ? carry[v] : up;
        carry[v] = last;
      }
      if (fin) {
        double r[5];
#pragma unroll
        for (int v = 0; v < 5; ++v) r[v] = Fl[v] - F[v];
        finalize_cell<EPI>(g, a, hidx, ridx, in, r, step, red, f - 1, i1, i2);
      }
      gf += 32;
      f += 32;
      while (f >= nf) {
        f -= nf;
        ++row;
      }
    }
  }
  if (EPI) {
    if (a.reduce) red_commit(red, a.red);
  }
}

// ---------------------------------------------------------------------------
// contiguous sweep, production form ("rows"): a warp owns groups of 32 rows.  Per row it walks
// faces 1..nA in full 32-lane iterations (cell f-1 is finalised by the lane that computes face f;
// lane 0 takes the previous iteration's last flux as carry), so every lane always has work; the
// 32 row-opening faces f=0 of a group are computed first, one per lane.
// Windows: the 37 cells [32 it - 2, 32 it + 34] of the row that one iteration touches are staged
// into a per-warp shared-memory buffer ONE ITERATION AHEAD -- by a TMA tensor copy
// (cp.async.bulk.tensor, box = 40 cells x 5 variables, completion on a per-buffer mbarrier) or,
// when the buffer pitch is not 16-byte aligned (odd extents), by per-lane cp.async -- so the
// DRAM latency of the next window hides behind the ~800 FP64 instructions of the current face.
// ---------------------------------------------------------------------------
constexpr int kWinSlots = 40;                  // cells per staged window (37 used)
constexpr int kWinBytes = 5 * kWinSlots * 8;   // 1600 B moved per TMA op
constexpr int kWinStride = 1664;               // buffer pitch, multiple of 128 B (TMA destination alignment)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 4-D tiled TMA load: coordinates (c0 = contiguous cell index, c1, c2, c3 = variable)
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// out-of-line flux from a strided global window (rare paths only)
template <int A, int RECON, int RIEMANN>
__device__ __noinline__ void face_flux_from_global(const double* base, long long vst, long long sA, double gamma,
                                                   double (&F)[5], int lim, const FluxLimArgs fl) {
  double w[5][6];
#pragma unroll
  for (int v = 0; v < 5; ++v)
#pragma unroll
    for (int k = 0; k < 6; ++k) w[v][k] = base[v * vst + k * sA];
  face_flux<A, RECON, RIEMANN>(w, gamma, F, lim, fl);
}

struct RowsArgs {
  int iters_per_row;      // ceil(nA / 32)
  int group_rows;         // rows per warp work item (<= 32)
  int shift;              // window slot 0 holds cell 32*it - shift; 2 or 3 so that the TMA start
                          // coordinate (cA_off + 32*it - shift) is even: UTMALDG traps unless the
                          // innermost coordinate * element size is a multiple of 16 B (measured on B200)
  int c1_off, c2_off;     // halo offsets of the two transverse roles in TMA coordinates
  int cA_off;             // halo offset of the sweep axis
  int tma_dim1_is_role;   // which role (1 or 2) is TMA dimension 1 (the faster transverse axis): always role 2
};

template <int A, int RECON, int RIEMANN, int EPI, int USE_TMA>
__global__ void __launch_bounds__(128, JXF_MIN_BLOCKS)
sweep_rows(const __grid_constant__ SweepGeom g, const __grid_constant__ SweepArgs a, const RowsArgs ra, const __grid_constant__ CUtensorMap tmap) {
  __shared__ alignas(128) unsigned char win_raw[4 * 2 * kWinStride];
  __shared__ alignas(8) uint64_t bars[4 * 2];
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const long long gwarp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long ngwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long nrows = (long long)g.n1 * g.n2;
  const int G = ra.group_rows;
  const long long ngroups = (nrows + G - 1) / G;
  const int ipr = ra.iters_per_row;
  const double step = (EPI ? (*a.dt) * a.dt_mult : 0.0);
  unsigned char* const win0 = win_raw + wid * 2 * kWinStride;      // this warp's two window buffers
  uint64_t* const bar0 = &bars[wid * 2];
  uint32_t phase_bits = 0u;                                        // bit b = parity to wait for on buffer b
  if (USE_TMA) {
    if (lane == 0) {
      mbar_init(bar0, 1);
      mbar_init(bar0 + 1, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
  }
  Red red;
  red.init();

  for (long long group = gwarp; group < ngroups; group += ngwarps) {
    const long long row0 = group * G;
    const int nr = (int)min((long long)G, nrows - row0);
    // ---- the row-opening faces f = 0 of this group, one row per lane (direct strided loads; the flux
    // code is called out of line here so the hot loop below holds the only inlined copy) -----------
    double F0[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    if (lane < nr) {
      const long long row = row0 + lane;
      const int k1 = (int)(row / g.n2);
      const int k2 = (int)(row - (long long)k1 * g.n2);
      face_flux_from_global<A, RECON, RIEMANN>(a.prims + k1 * g.s1 + k2 * g.s2 - 3 * g.sA, g.vst, g.sA, a.gamma, F0, a.limiter, a.fl);
    }
    // ---- main sequence: (row r, iteration it), windows staged one step ahead ---------------------
    // All index state is carried incrementally in 32-bit registers (no divisions in the loop):
    // (r, it, i1, i2) for the iteration being computed, (rn, itn, i1n, i2n) for the one being staged.
    const int total = nr * ipr;
    const int i1_0 = (int)(row0 / g.n2);
    const int i2_0 = (int)(row0 - (long long)i1_0 * g.n2);
    auto issue = [&](int b, int itn, int i1n, int i2n) {
      double* const wb = reinterpret_cast<double*>(win0 + b * kWinStride);
      if (USE_TMA) {
        if (lane == 0) {
          mbar_expect_tx(bar0 + b, kWinBytes);
          // TMA dims: (contiguous sweep axis, faster transverse (role 2), slower transverse (role 1), variable);
          // cells past the end of the row are zero-filled by the TMA unit, no predication needed
          tma_load_4d(wb, &tmap, bar0 + b, ra.cA_off + 32 * itn - ra.shift, ra.c2_off + i2n, ra.c1_off + i1n, 0);
        }
      } else {
        const double* src = a.prims + i1n * g.s1 + i2n * g.s2 + (long long)(32 * itn - ra.shift) * g.sA;
        const int cmax = g.nA + 2 - (32 * itn - ra.shift);  // slots holding cells <= nA+2 are valid
#pragma unroll
        for (int v = 0; v < 5; ++v) {
          if (lane <= cmax) cp_async8(wb + v * kWinSlots + lane, src + v * g.vst + lane);
          if (lane < 6 && 32 + lane <= cmax) cp_async8(wb + v * kWinSlots + 32 + lane, src + v * g.vst + 32 + lane);
        }
        cp_async_commit();
      }
    };
    issue(0, 0, i1_0, i2_0);
    double carry[5];
    int r = 0, it = 0, i1 = i1_0, i2 = i2_0;
    int itn = 0, i1n = i1_0, i2n = i2_0;
    long long col_h = i1 * g.s1 + i2 * g.s2;
    long long col_r = i1 * g.r1 + i2 * g.r2;
    for (int j = 0; j < total; ++j) {
      const int b = j & 1;
      // advance the staged-iteration state and stage it
      if (++itn == ipr) {
        itn = 0;
        if (++i2n == g.n2) {
          i2n = 0;
          ++i1n;
        }
      }
      if (j + 1 < total) issue(b ^ 1, itn, i1n, i2n);
      const int f = 1 + 32 * it + lane;
      const bool act = f <= g.nA;
      const long long hidx = col_h + (long long)(f - 1) * g.sA;
      const long long ridx = col_r + (long long)(f - 1) * g.rA;
      CellIn<EPI> in;
      if (act) load_cell_in<EPI>(g, a, hidx, ridx, in);
      if (it == 0) {
#pragma unroll
        for (int v = 0; v < 5; ++v) carry[v] = __shfl_sync(0xffffffffu, F0[v], r);
      }
      // wait for this iteration's window
      const double* const wb = reinterpret_cast<const double*>(win0 + b * kWinStride);
      if (USE_TMA) {
        mbar_wait(bar0 + b, (phase_bits >> b) & 1u);
        phase_bits ^= (1u << b);
      } else {
        if (j + 1 < total) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncwarp();
      }
      double F[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
      if (act) {
        double w[5][6];
        const double* wl = wb + (ra.shift - 2) + lane;
#pragma unroll
        for (int v = 0; v < 5; ++v)
#pragma unroll
          for (int k = 0; k < 6; ++k) w[v][k] = wl[v * kWinSlots + k];
        face_flux<A, RECON, RIEMANN>(w, a.gamma, F, a.limiter, a.fl);
      }
      double Fl[5];
#pragma unroll
      for (int v = 0; v < 5; ++v) {
        const double up = __shfl_up_sync(0xffffffffu, F[v], 1);
        const double last = __shfl_sync(0xffffffffu, F[v], 31);
        Fl[v] = (lane == 0) ? carry[v] : up;
        carry[v] = last;
      }
      if (act) {
        double rr[5];
#pragma unroll
        for (int v = 0; v < 5; ++v) rr[v] = Fl[v] - F[v];
        finalize_cell<EPI>(g, a, hidx, ridx, in, rr, step, red, f - 1, i1, i2);
      }
      __syncwarp();          // all lanes are done with win[b] before it is refilled two iterations later
      if (++it == ipr) {
        it = 0;
        ++r;
        if (++i2 == g.n2) {
          i2 = 0;
          ++i1;
        }
        col_h = i1 * g.s1 + i2 * g.s2;
        col_r = i1 * g.r1 + i2 * g.r2;
      }
    }
  }
  if (EPI) {
    if (a.reduce) red_commit(red, a.red);
  }
}

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
    const int i2 = (int)(q % n2);
    const long long q1 = q / n2;
    const int i1 = (int)(q1 % n1);
    const int l = (int)(q1 / n1);   // halo layer in increasing buffer index
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
  const long long total = (long long)g.nh * n1 * n2;
  const int nh = g.nh, ext = g.ext[ax];
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total;
       q += (long long)gridDim.x * blockDim.x) {
    // slab order: layers slowest, except for faces of the CONTIGUOUS axis, whose nh layers are adjacent in memory:
    // there the layer index runs fastest so that a row's nh cells are one 8 nh-byte access (both ends of an
    // exchange run this kernel, so the order is private to it)
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
    const long long tr = (long long)(i1 + a.lo1) * g.st[t1] + (long long)(i2 + a.lo2) * g.st[t2];
    if (!a.unpack) {
      const int src = hi ? (ext - 2 * nh + l) : (nh + l);   // interior layers adjacent to the face
      const long long is = tr + (long long)src * g.st[ax];
#pragma unroll
      for (int v = 0; v < 5; ++v) a.slab[q + v * total] = a.prims[is + v * g.vst];
    } else {
      const int dst = hi ? (ext - nh + l) : l;
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

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

static int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(JXF_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return JXF_OK;
}

struct jxf_solver {
  jxf_config cfg;
  Geom g;
  int active[3];
  int n_active;
  int active_mask;
  int lane_axis;      // contiguous active axis
  int order[3];       // order[k] = axis of the k-th sweep of a stage; the LAST one carries the fused epilogue
  const double* dt_bound;   // jxf_bind_timestep: time step for the flux limiter outside jxf_stage
  int num_sms;
  int stages;
  double dt_mult[4];
  double blend[4][2];
  // TMA descriptors of the primitive buffers seen so far (keyed by base pointer)
  bool force_rows;     // JXF_FORCE_ROWS=1: use the rows kernel on small grids too (tests)
  bool tma_ok;
  bool no_march;       // -DJXF_WITH_STRIDED builds, JXF_NO_MARCH=1: register-window strided kernel (A/B only)
  int n_maps;
  const void* map_ptr[8];
  CUtensorMap map[8];
  // launch accounting / optional per-kernel event timing (jxf_profile_*)
  long long launches[JXF_PROFILE_KINDS];
  int prof_on;
  int prof_n;
  int prof_cap;
  cudaEvent_t* prof_start;
  cudaEvent_t* prof_stop;
  unsigned char* prof_kind;
};

struct ProfScope {
  jxf_solver* s;
  cudaStream_t st;
  int slot;
  ProfScope(const jxf_solver* cs, int kind, cudaStream_t stream) : s(const_cast<jxf_solver*>(cs)), st(stream), slot(-1) {
    s->launches[kind]++;
    if (s->prof_on && s->prof_n < s->prof_cap) {
      slot = s->prof_n++;
      s->prof_kind[slot] = (unsigned char)kind;
      cudaEventRecord(s->prof_start[slot], st);
    }
  }
  ~ProfScope() {
    if (slot >= 0) cudaEventRecord(s->prof_stop[slot], st);
  }
};

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
static const CUtensorMap* get_rows_map(jxf_solver* s, const double* base) {
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

// ---------------------------------------------------------------------------
// sweep dispatch
// ---------------------------------------------------------------------------
static void set_role_bcs(SweepGeom& sg, const SweepArgs& a) {
  sg.bcA_hi = a.bc[2 * sg.axA]; sg.bcA_lo = a.bc[2 * sg.axA + 1];
  sg.bc1_hi = a.bc[2 * sg.ax1]; sg.bc1_lo = a.bc[2 * sg.ax1 + 1];
  sg.bc2_hi = a.bc[2 * sg.ax2]; sg.bc2_lo = a.bc[2 * sg.ax2 + 1];
}

template <int A, int RECON, int RIEMANN, int EPI>
static int launch_sweep(const jxf_solver* s, SweepArgs a, cudaStream_t st) {
  const Geom& g = s->g;
  const int resident = s->num_sms * JXF_MIN_BLOCKS;   // CTAs of 128 threads resident on the device
  // fold the interior origin into the base pointers
  const long long h0 = g.off[0] * g.st[0] + g.off[1] * g.st[1] + g.off[2] * g.st[2];
  a.prims += h0;
  if (a.cons_in) a.cons_in += h0;
  if (a.cons_n) a.cons_n += h0;
  if (a.cons_out) a.cons_out += h0;
  if (a.prims_out) a.prims_out += h0;
  SweepGeom sg;
  sg.axA = A;
  sg.nA = g.n[A];
  sg.sA = g.st[A];
  sg.rA = g.rst[A];
  sg.vst = g.vst;
  sg.rvst = g.rvst;
  const int T1 = (A == 0) ? 1 : 0;       // slower transverse axis
  const int T2 = (A == 2) ? 1 : 2;       // faster transverse axis
  if (A != s->lane_axis) {
    // lanes along the contiguous axis C; the other transverse axis O is the slow one
    const int C = s->lane_axis;
    const int O = 3 - A - C;
    sg.ax1 = O; sg.n1 = g.n[O]; sg.s1 = g.st[O]; sg.r1 = g.rst[O];
    sg.ax2 = C; sg.n2 = g.n[C]; sg.s2 = g.st[C]; sg.r2 = g.rst[C];
    const long long plane = (long long)sg.n1 * sg.n2;
    const int bx = (int)((plane + 127) / 128);
    const int resident = s->num_sms * (s->no_march ? JXF_MIN_BLOCKS : JXF_MARCH_BLOCKS);
    // chunks along A: every chunk costs one redundant face (+ a 5-plane prologue), while few CTAs per
    // resident slot leave a partial last wave; pick the chunk count that minimises
    // (1 + 1.5/chunk_len) * ceil(waves)/waves over chunk lengths >= 16 cells
    if (a.range_hi <= a.range_lo) { a.range_lo = 0; a.range_hi = g.n[A]; }
    const int nr = a.range_hi - a.range_lo;
    int chunks = 1;
    double best = 1e30;
    const int max_chunks = std::max(1, std::min(nr / 16, 65535));
    for (int c = 1; c <= max_chunks; ++c) {
      const int len = (nr + c - 1) / c;
      const int cc = (nr + len - 1) / len;
      const double waves = (double)bx * cc / resident;
      const double cost = (1.0 + 1.5 / len) * (waves <= 1.0 ? 1.0 / waves : std::ceil(waves) / waves);
      if (cost < best - 1e-12) { best = cost; chunks = cc; }
    }
    a.chunk_len = (nr + chunks - 1) / chunks;
    chunks = (nr + a.chunk_len - 1) / a.chunk_len;
    dim3 grid(bx, chunks);
    set_role_bcs(sg, a);
    ProfScope prof(s, A + 3 * EPI, st);
#ifdef JXF_WITH_STRIDED
    if (s->no_march) {
      sweep_strided<A, RECON, RIEMANN, EPI><<<grid, 128, 0, st>>>(sg, a);
      return check_launch("sweep_strided");
    }
#endif
    sweep_march<A, RECON, RIEMANN, EPI><<<grid, 128, 0, st>>>(sg, a);
  } else {
    sg.ax1 = T1; sg.n1 = g.n[T1]; sg.s1 = g.st[T1]; sg.r1 = g.rst[T1];
    sg.ax2 = T2; sg.n2 = g.n[T2]; sg.s2 = g.st[T2]; sg.r2 = g.rst[T2];
    set_role_bcs(sg, a);
    const long long rows = (long long)sg.n1 * sg.n2;
    const int nf = g.n[A] + 1;
    const long long total = rows * nf;
#if JXF_ROWS_KERNEL
    // production form whenever groups of >= 4 rows give every resident warp several work items
    const long long warps_resident = 4LL * resident;
    if ((s->force_rows || rows / 4 >= warps_resident * 2) && g.n[A] >= 32) {
      RowsArgs ra;
      ra.iters_per_row = (g.n[A] + 31) / 32;
      // one group per warp, 4 warps per CTA, many more CTAs than resident slots: the hardware block
      // scheduler balances the tail (a static groups-per-warp split left ~8 % of the warps idle at the end)
      int G = 8;
      while (G > 4 && rows / G < warps_resident * 16) G >>= 1;
      ra.group_rows = G;
      ra.shift = ((g.off[A] - 2) & 1) ? 3 : 2;
      ra.cA_off = g.off[A];
      ra.c1_off = g.off[T1];
      ra.c2_off = g.off[T2];
      ra.tma_dim1_is_role = 2;
      const long long groups = (rows + G - 1) / G;
      const long long blocks = std::min<long long>((groups + 3) / 4, 1LL << 30);
      const CUtensorMap* map = get_rows_map(const_cast<jxf_solver*>(s), a.prims - h0);
      ProfScope prof(s, A + 3 * EPI, st);
      if (map) {
        sweep_rows<A, RECON, RIEMANN, EPI, 1><<<(unsigned)blocks, 128, 0, st>>>(sg, a, ra, *map);
      } else {
        CUtensorMap dummy;
        memset(&dummy, 0, sizeof(dummy));
        sweep_rows<A, RECON, RIEMANN, EPI, 0><<<(unsigned)blocks, 128, 0, st>>>(sg, a, ra, dummy);
      }
      return check_launch("sweep_rows");
    }
#endif
    const long long target_warps = 4LL * resident * 4;   // ~4 waves of warps
    long long span;
    if (rows >= target_warps) {
      span = (rows / target_warps) * nf;                 // whole rows per range, no carry-in face
      span = std::min<long long>(span, 64LL * nf);
    } else {
      span = std::max<long long>(31, ((total / target_warps) / 32) * 32 - 1);
      span = std::min<long long>(span, 32LL * 256 - 1);
    }
    if (span > 0x7fffffff) span = 0x7fffffff;
    a.span = (int)span;
    const long long nranges = (total + span - 1) / span;
    const long long blocks = std::min<long long>((nranges + 3) / 4, (long long)resident * 4);
    ProfScope prof(s, A + 3 * EPI, st);
    sweep_contig<A, RECON, RIEMANN, EPI><<<(unsigned)std::max<long long>(1, blocks), 128, 0, st>>>(sg, a, total);
  }
  return check_launch("sweep");
}

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
  switch (recon_template_of(s)) {
    case 0: return dispatch_riemann<A, 0>(s, a, epi, st);
    case 1: return dispatch_riemann<A, 1>(s, a, epi, st);
    case 2: return dispatch_riemann<A, 2>(s, a, epi, st);
    case 3: return dispatch_riemann<A, 3>(s, a, epi, st);
    case 4: return dispatch_riemann<A, 4>(s, a, epi, st);
    default: return dispatch_riemann<A, 5>(s, a, epi, st);
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

extern "C" int jxf_halo_fill(jxf_handle h, double* prims, double* cons, void* stream) {
  if (!h || !prims || !cons) return fail(JXF_ERR_BAD_ARG, "jxf_halo_fill: null argument");
  HaloArgs a;
  a.prims = prims;
  a.cons = cons;
  a.gamma = h->cfg.gamma;
  long long maxcells = 0;
  for (int f = 0; f < 6; ++f) {
    a.bc[f] = h->cfg.bc[f];
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
    ProfScope prof(h, JXF_PROFILE_HALO, (cudaStream_t)stream);
    halo_fill_kernel<<<dim3(bx, 6), 128, 0, (cudaStream_t)stream>>>(h->g, a);
    int rc = check_launch("halo_fill");
    if (rc) return rc;
  }
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
      rc = dispatch_axis(h, axis, a, 1, (cudaStream_t)stream);
    }
    if (rc) return rc;
  }
  // the next stage's dissipative stencils read edge halos (halo_manager.py:119-129)
  if (diss && fill_halo) return jxf_halo_fill_edges(h, prims_out, cons_out, stream);
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

extern "C" int64_t jxf_face_slab_elems_ext(jxf_handle h, int face, int ext_mask) {
  if (!h || face < 0 || face > 5) return -1;
  int lo1, n1, lo2, n2;
  slab_ranges(h, face, ext_mask, lo1, n1, lo2, n2);
  return 5LL * h->g.nh * n1 * n2;
}

extern "C" int64_t jxf_face_slab_elems(jxf_handle h, int face) { return jxf_face_slab_elems_ext(h, face, 0); }

static int face_slab(jxf_handle h, int face, int ext_mask, double* prims, double* cons, double* slab, int unpack, void* stream) {
  if (!h || !prims || !slab || (unpack && !cons)) return fail(JXF_ERR_BAD_ARG, "jxf_(un)pack_face: null argument");
  if (face < 0 || face > 5 || h->g.n[face >> 1] <= 1) return fail(JXF_ERR_BAD_ARG, "jxf_(un)pack_face: face %d not active", face);
  if (ext_mask < 0 || ext_mask > 15) return fail(JXF_ERR_BAD_ARG, "jxf_(un)pack_face: ext_mask %d", ext_mask);
  FaceArgs a;
  a.prims = prims;
  a.cons = cons;
  a.slab = slab;
  a.gamma = h->cfg.gamma;
  a.face = face;
  a.unpack = unpack;
  slab_ranges(h, face, ext_mask, a.lo1, a.n1, a.lo2, a.n2);
  const long long total = (long long)h->g.nh * a.n1 * a.n2;
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
