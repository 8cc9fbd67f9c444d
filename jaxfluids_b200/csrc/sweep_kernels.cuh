// sweep_kernels.cuh -- the convective sweep kernels (one family per sweep direction kind) and the structures they take.
// Included by jxf_b200.cu (C ABI, dispatch) and by sweep_inst.cu (one translation unit per (axis, RECON) pair, so that
// the ~330 kernel instantiations compile in parallel).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

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

// Per-face boundary DATA on the device (jxf_set_face_data): what NEUMANN / SIMPLE_INFLOW / SIMPLE_OUTFLOW faces,
// DIRICHLET faces with space-dependent primitives_callable, WALL faces with a space-dependent velocity and faces
// with several types prescribe on top of the face's base rule (halos/outer/material.py:473-520, :732-866, :966-1050).
// The base rule (bc[face]: ZEROGRADIENT copy of the last interior cell, the WALL / SYMMETRY mirror, ...) produces the
// halo primitives; then, per variable v with op = (ops >> 2 v) & 3:  1: p_v = data_v,  2: p_v = p_v + data_v
// (NEUMANN increment (value * sign) * dx; WALL 2 u_wall on top of -u_mirror), 0: keep.  data: (5, n1, n2) over the
// face's transverse interior cells, the same for every halo layer (the reference expands the callable's values along
// the face normal); mask (n1, n2) or null: apply only where mask != 0 (one type of a multi-type face).
struct FaceData {
  const double* data[6];
  const unsigned char* mask[6];
  int ops[6];
};

__device__ __forceinline__ void apply_face_data(const FaceData& fd, int face, long long tidx, long long tcount, double (&p)[5]) {
  const int ops = fd.ops[face];
  if (ops == 0) return;
  if (fd.mask[face] && fd.mask[face][tidx] == 0) return;
  const double* d = fd.data[face] + tidx;
#pragma unroll
  for (int v = 0; v < 5; ++v) {
    const int op = (ops >> (2 * v)) & 3;
    if (op == 1) p[v] = d[v * tcount];
    else if (op == 2) p[v] = p[v] + d[v * tcount];
  }
}

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
  // slab launches (jxf_stage_inplace): a sweep along y or z restricted to the x planes [sub_lo, sub_lo + sub_n);
  // rvst_slab: variable stride of the slab-sized rhs accumulator (0: the full-size one)
  int sub_lo, sub_n;
  long long rvst_slab;
  int inplace;             // EPI: prims_out aliases prims (rows kernel: the next window must have landed before a store)
  // EPI + fuse_halo, multi-GPU: output buffers of the block across each NEIGHBOR face, mapped into this process (CUDA
  // IPC over NVLink), or null.  The thread that produces a cell within nh of such a face stores the cell's image -- the
  // neighbour's halo cell -- straight into the neighbour's memory (halo_images_axis), instead of a pack / send / unpack
  // round through staging slabs.  Pointers carry the same interior-origin offset as prims_out / cons_out.
  double* peer_prims[6];
  double* peer_cons[6];
  FaceData face_data;      // EPI + fuse_halo: per-face boundary data (device pointers), used when has_face_data
  int has_face_data;
  int n_phys[3];           // interior cells per PHYSICAL axis (transverse indexing of face_data)
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
  int i1_base, n1_full;      // slab launches (jxf_stage_inplace): role 1 covers cells [i1_base, i1_base + n1) of n1_full
  // cells with iA < nearA_lo or iA >= nearA_hi owe halo images to the faces of role A: nh / nA - nh, or 0 / nA when
  // those faces are filled by a separate launch after the sweep (plan.cuh: rows kernel, every row END is such a cell)
  int nearA_lo, nearA_hi;
  // faces of role A whose images finalize_cell writes INLINE (no call): 0 = no, 1 = SYMMETRY, 2 = PERIODIC -- faces without
  // boundary data / peer stores only; the generic path then sees these faces as INACTIVE (plan.cuh)
  int leanA_lo, leanA_hi;
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

// a * b + c of the cell update: one fused multiply-add in the production evaluation; separate multiply and add -- the
// reference's operations -- in JXF_REFERENCE_ORDER builds (compiled with -fmad=false for the strict-norm parity test)
#ifdef JXF_REFERENCE_ORDER
#define JXF_MADD(a, b, c) ((a) * (b) + (c))
#else
#define JXF_MADD(a, b, c) fma((a), (b), (c))
#endif

// EPI template parameter of the sweep kernels:
//   0      no epilogue: the axis contribution goes to the rhs accumulator;
//   1      epilogue with RUN-TIME flags (has_prev, blend, reduce, volume_force, fuse_halo of SweepArgs);
//   2..5   "fast" epilogue: has_prev = 1, volume_force = 0 and the flags (EPI - 2) = blend | reduce << 1 are COMPILE-TIME,
//          so the hot loop carries no flag loads / branches for them (launch_sweep picks the instantiation from the
//          run-time flags; same arithmetic).  fuse_halo stays a run-time flag (its call sits in a cold branch).
template <int EPI>
struct EpiFlags {
  static constexpr bool kStatic = EPI >= 2;
  static constexpr bool kBlend = ((EPI - 2) & 1) != 0;
  static constexpr bool kReduce = ((EPI - 2) & 2) != 0;
  __device__ static __forceinline__ bool has_prev(const SweepArgs& a) { return kStatic ? true : a.has_prev != 0; }
  __device__ static __forceinline__ bool blend(const SweepArgs& a) { return kStatic ? kBlend : a.blend != 0; }
  __device__ static __forceinline__ bool reduce(const SweepArgs& a) { return kStatic ? kReduce : a.reduce != 0; }
  __device__ static __forceinline__ bool halo(const SweepArgs& a) { return a.fuse_halo != 0; }
  __device__ static __forceinline__ bool force(const SweepArgs& a) { return kStatic ? false : a.volume_force != 0; }
};

// operands of the cell update that come from memory; loaded EARLY (before the flux arithmetic of
// the iteration) so their latency hides behind ~700 FP64 instructions
template <int EPI>
struct CellIn {
  double rhs[5];   // EPI=0: accumulate target (if accumulate); EPI>=1: earlier axes' sum (if has_prev)
  double U[5];     // EPI>=1
  double Un[5];    // EPI>=1, blend
};

template <int EPI>
__device__ __forceinline__ void load_cell_in(const SweepGeom& g, const SweepArgs& a, long long hidx, long long ridx,
                                             CellIn<EPI>& in) {
  using E = EpiFlags<EPI>;
  if (EPI == 0) {
    if (a.accumulate) {
#pragma unroll
      for (int v = 0; v < 5; ++v) in.rhs[v] = a.rhs[ridx + v * g.rvst];
    }
  } else {
#pragma unroll
    for (int v = 0; v < 5; ++v) in.U[v] = a.cons_in[hidx + v * g.vst];
    if (E::has_prev(a)) {
#pragma unroll
      for (int v = 0; v < 5; ++v) in.rhs[v] = a.rhs[ridx + v * g.rvst];
    }
    if (E::blend(a)) {
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
//
// Cost matters here: in the rows kernel every row end is such a cell, so one warp iteration in six takes this path.
// The images of ONE face are one value written `count` times (count = 1, or nh for ZEROGRADIENT / DIRICHLET), so a
// face reduces to a small descriptor and ONE out-of-line writer (write_images) holds the only copy of the image
// arithmetic (velocity flip / wall reflection, boundary data, conservatives with the IEEE division of the
// halo-fill kernel, ten stores).
struct HaloOut {
  double* prims;
  double* cons;
  long long vst;
  double gamma;
  int nh;
};

// wall = nullptr: copy, negating velocity component flip_var (1..3; -1 = none).  wall != nullptr: no-slip wall
// moving with (u, v, w) = wall[0..2]: every velocity component becomes 2 u_wall - u (halos/outer/material.py:510-512).
// fixed != nullptr: the face's DIRICHLET constants replace the cell's primitives.  fd / face / tidx / tcount: the
// face's device-resident boundary data, applied on top (apply_face_data).  The image goes to dst, dst + dinc, ...
static __device__ __noinline__ void write_images(double* prims, double* cons, long long dst, long long dinc, int count,
                                                 long long vst, double gamma, double p0, double p1, double p2, double p3,
                                                 double p4, int flip_var, const double* wall, const double* fixed,
                                                 const FaceData* fd, int face, long long tidx, long long tcount) {
  double q[5] = {p0, p1, p2, p3, p4};
  if (fixed) {
#pragma unroll
    for (int v = 0; v < 5; ++v) q[v] = fixed[v];
  }
  if (wall) {
#pragma unroll
    for (int v = 1; v < 4; ++v) q[v] = 2 * wall[v - 1] - q[v];
  } else {
#pragma unroll
    for (int v = 1; v < 4; ++v) q[v] = (v == flip_var) ? q[v] * -1.0 : q[v];
  }
  if (fd) apply_face_data(*fd, face, tidx, tcount, q);
  double c[5];
  cons_from_prims(q, gamma, c);
#pragma unroll 1
  for (int l = 0; l < count; ++l, dst += dinc) {
#pragma unroll
    for (int v = 0; v < 5; ++v) {
      prims[dst + v * vst] = q[v];
      cons[dst + v * vst] = c[v];
    }
  }
}

// the images one cell owes to ONE face of a role axis (high: east / north / top).  bc: the face's rule; i / n / stride:
// the cell's index, the cell count and the element stride along the axis; tidx / tcount: face-data indexing.
__device__ __forceinline__ void face_images(const HaloOut& own, const SweepArgs& a, const FaceData* fd, int bc, bool high,
                                            int ax, int n, int i, long long stride, long long hidx, const double (&p)[5],
                                            long long tidx, long long tcount) {
  const int nh = own.nh;
  const int face = 2 * ax + (high ? 0 : 1);
  const bool at_face = high ? (i >= n - nh) : (i < nh);       // within nh of THIS face
  const bool at_opposite = high ? (i < nh) : (i >= n - nh);   // ... of the other face of the axis
  const long long out = high ? stride : -stride;              // one cell outwards through this face
  double* bp = own.prims;
  double* bq = own.cons;
  long long dst = 0, dinc = 0;
  int count = 0, flip = -1;
  const double* wall = nullptr;
  const double* fixed = nullptr;
  bool with_fd = true;
  if (bc == JXF_BC_SYMMETRY || bc == JXF_BC_WALL) {           // mirror image across the face
    if (at_face) {
      count = 1;
      dst = hidx + (long long)(high ? 2 * (n - i) - 1 : -1 - 2 * i) * stride;
      if (bc == JXF_BC_SYMMETRY) flip = 1 + ax; else wall = a.wall[face];
    }
  } else if (bc == JXF_BC_PERIODIC) {                         // this face's halo = the cells next to the OTHER face
    if (at_opposite) { count = 1; dst = hidx + (long long)n * out; with_fd = false; }
  } else if (bc == JXF_BC_ZEROGRADIENT || bc == JXF_BC_DIRICHLET) {   // the boundary-adjacent cell writes all nh layers
    if (high ? (i == n - 1) : (i == 0)) {
      count = nh; dst = hidx + out; dinc = out;
      if (bc == JXF_BC_DIRICHLET) fixed = a.dirichlet[face];
    }
  } else if (bc == JXF_BC_NEIGHBOR) {
    // shared with another block: the nh cells next to the face are the neighbour's halo cells beyond ITS opposite face --
    // the index arithmetic of a periodic image, into the neighbour's (peer-mapped) buffers
    if (at_face && a.peer_prims[face]) {
      count = 1; bp = a.peer_prims[face]; bq = a.peer_cons[face]; dst = hidx - (long long)n * out; with_fd = false;
    }
  }
  if (count)
    write_images(bp, bq, dst, dinc, count, own.vst, own.gamma, p[0], p[1], p[2], p[3], p[4], flip, wall, fixed,
                 with_fd ? fd : nullptr, face, tidx, tcount);
}

// OUT OF LINE on purpose: executed only by the thin shell of boundary-adjacent cells; keeping it out of
// the sweep loop keeps the hot loop short (instruction cache) and its register allocation free of this code
static __device__ __noinline__ void halo_images_cell(const SweepGeom& g, const SweepArgs& a, long long hidx, double p0, double p1,
                                              double p2, double p3, double p4, int iA, int i1, int i2);

template <int EPI>
__device__ __forceinline__ void finalize_cell(const SweepGeom& g, const SweepArgs& a, long long hidx, long long ridx,
                                              const CellIn<EPI>& in, const double (&r)[5], double step, Red& red,
                                              int iA, int i1, int i2) {
  using E = EpiFlags<EPI>;
  // r = F_{i-1/2} - F_{i+1/2}; the axis contribution (1/dx) r (space_solver.py:597-599) is added to the
  // earlier axes' sum with one fused multiply-add
  if (EPI == 0) {
#pragma unroll
    for (int v = 0; v < 5; ++v) a.rhs[ridx + v * g.rvst] = a.accumulate ? JXF_MADD(a.inv_dx, r[v], in.rhs[v]) : a.inv_dx * r[v];
  } else {
    double U[5];
#pragma unroll
    for (int v = 0; v < 5; ++v) {
      double tot = E::has_prev(a) ? JXF_MADD(a.inv_dx, r[v], in.rhs[v]) : a.inv_dx * r[v];
      if (E::force(a)) {     // space_solver.py:378-384 with the einsums of source_term_solver.py:180-182
        if (v >= 1 && v <= 3) tot += a.gravity[v - 1] * in.U[0];
        if (v == 4) tot += (a.gravity[0] * in.U[1] + a.gravity[1] * in.U[2]) + a.gravity[2] * in.U[3];
      }
      double u = in.U[v];
      if (E::blend(a)) u = a.ca * u + a.cb * in.Un[v];
      U[v] = u + step * tot;
    }
    double p[5];
    prims_from_cons(U, a.gamma, p);
#pragma unroll
    for (int v = 0; v < 5; ++v) {
      a.cons_out[hidx + v * g.vst] = U[v];
      a.prims_out[hidx + v * g.vst] = p[v];
    }
    if (E::reduce(a)) {
#ifndef JXF_REFERENCE_ORDER
      if (E::kStatic) red.add_cell_fast(p, a.gamma, a.active_mask); else
#endif
      red.add_cell(p, a.gamma, a.active_mask);
    }
    if (E::halo(a)) {
      // mirror / periodic images across the faces of the sweep's own axis: in the rows kernel these are the row ends (one
      // warp iteration in eight, 5 active lanes).  The image is the cell with one velocity negated, or the cell itself:
      // destination and flip are formed here from kernel constants and handed to the one image writer, without the
      // descriptor logic of halo_images_cell (which reads the launch's geometry through generic loads)
      if ((g.leanA_lo | g.leanA_hi) && ((iA < a.nh) | (iA >= g.nA - a.nh))) {
#pragma unroll 1
        for (int end = 0; end < 2; ++end) {
          const bool here = end ? (iA >= g.nA - a.nh) : (iA < a.nh);
          // the cells at this end feed THIS end's face when it mirrors, the OTHER end's face when the axis is periodic
          const int mine = end ? g.leanA_hi : g.leanA_lo, other = end ? g.leanA_lo : g.leanA_hi;
          if (here & ((mine == 1) | (other == 2))) {
            const long long off = (mine == 1) ? (end ? 2 * (g.nA - iA) - 1 : -1 - 2 * iA) : (end ? -g.nA : g.nA);
            write_images(a.prims_out, a.cons_out, hidx + off * g.sA, 0, 1, g.vst, a.gamma, p[0], p[1], p[2], p[3], p[4],
                         (mine == 1) ? 1 + g.axA : -1, nullptr, nullptr, nullptr, 0, 0, 0);
          }
        }
      }
      // boundary-adjacent cells only (a thin shell); warp-divergent by construction
      const int j1 = i1 + g.i1_base;      // global index along role 1 (slab launches)
      const bool near = (iA < g.nearA_lo) | (iA >= g.nearA_hi) | (j1 < a.nh) | (j1 >= g.n1_full - a.nh) | (i2 < a.nh) |
                        (i2 >= g.n2 - a.nh);
      if (near) halo_images_cell(g, a, hidx, p[0], p[1], p[2], p[3], p[4], iA, j1, i2);
    }
  }
}

static __device__ __noinline__ void halo_images_cell(const SweepGeom& g, const SweepArgs& a, long long hidx, double p0, double p1,
                                              double p2, double p3, double p4, int iA, int i1, int i2) {
  const HaloOut h{a.prims_out, a.cons_out, g.vst, a.gamma, a.nh};
  const double p[5] = {p0, p1, p2, p3, p4};
  const FaceData* fd = a.has_face_data ? &a.face_data : nullptr;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int n = (r == 0) ? g.nA : (r == 1) ? g.n1_full : g.n2;
    const int i = (r == 0) ? iA : (r == 1) ? i1 : i2;
    // every rule takes its images from cells within nh of one of the axis' two faces
    if (n <= 1 || (i >= a.nh && i < n - a.nh)) continue;
    const int ax = (r == 0) ? g.axA : (r == 1) ? g.ax1 : g.ax2;
    const long long stride = (r == 0) ? g.sA : (r == 1) ? g.s1 : g.s2;
    const int bhi = (r == 0) ? g.bcA_hi : (r == 1) ? g.bc1_hi : g.bc2_hi;
    const int blo = (r == 0) ? g.bcA_lo : (r == 1) ? g.bc1_lo : g.bc2_lo;
    // face data is laid out over (t1, t2) = the two other PHYSICAL axes in increasing order, t2 fastest
    long long tidx = 0, tcnt = 0;
    if (fd) {
      const int t1 = (ax == 0) ? 1 : 0, t2 = (ax == 2) ? 1 : 2;
      const int j1 = (t1 == g.axA) ? iA : (t1 == g.ax1) ? i1 : i2;
      const int j2 = (t2 == g.axA) ? iA : (t2 == g.ax1) ? i1 : i2;
      tidx = (long long)j1 * a.n_phys[t2] + j2;
      tcnt = (long long)a.n_phys[t1] * a.n_phys[t2];
    }
    face_images(h, a, fd, blo, false, ax, n, i, stride, hidx, p, tidx, tcnt);
    face_images(h, a, fd, bhi, true, ax, n, i, stride, hidx, p, tidx, tcnt);
  }
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
    if (EpiFlags<EPI>::reduce(a)) red_commit(red, a.red);
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
    if (EpiFlags<EPI>::reduce(a)) red_commit(red, a.red);
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
    if (EpiFlags<EPI>::reduce(a)) red_commit(red, a.red);
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
  int tma_in;             // EPI: the cell inputs (U, U^n, rhs sum) are staged by TMA too (RowsInMaps), not loaded per lane
  int lead;               // ... their U / U^n boxes start `lead` (0 / 1) cells before the iteration's first cell
};

// Tensor maps of the epilogue's cell inputs (plan.cuh encode_rows_input_map): boxes of kInCells (U, U^n; halo'd
// conservative buffers) / 32 (rhs accumulator) cells x 5 variables, landing in the per-warp input buffers
#ifndef JXF_ROWS_TMA_IN
#define JXF_ROWS_TMA_IN 0
#endif
constexpr int kInCells = 34;                    // 32 cells + lead, rounded to a 16-byte multiple
constexpr int kInUBytes = 5 * kInCells * 8;     // 1360
constexpr int kInRBytes = 5 * 32 * 8;           // 1280
constexpr int kInUStride = 1408;                // 128-byte multiples (TMA destination alignment)
constexpr int kInStride = 2 * kInUStride + kInRBytes;   // 4096: U | U^n | rhs
struct RowsInMaps {
  CUtensorMap u, un, rhs;
};

#ifndef JXF_ROWS_LANE_CARRY
#define JXF_ROWS_LANE_CARRY 0
#endif
#ifndef JXF_ROWS_V1      // -DJXF_ROWS_V1: the round-1 form of the loop (A/B builds)
// Loop structure (round 2): rows and iterations are nested loops, so that the iteration body is ONE straight-line
// block -- no `it == 0` / `act` / `j + 1 < total` branches around the flux arithmetic, the TMA issue is predicated
// instead of branched, every lane evaluates its face (lanes past the end of the row work on the zero-filled tail of
// the window; only their loads / stores are predicated), and the left flux comes from ONE rotate-shuffle per
// variable (lane l <- lane l-1, lane 0 <- lane 31 = the carry of the next iteration).
template <int A, int RECON, int RIEMANN, int EPI, int USE_TMA>
__global__ void __launch_bounds__(128, JXF_MIN_BLOCKS)
sweep_rows(const __grid_constant__ SweepGeom g, const __grid_constant__ SweepArgs a, const RowsArgs ra, const __grid_constant__ CUtensorMap tmap,
           const __grid_constant__ RowsInMaps im) {
  __shared__ alignas(128) unsigned char win_raw[4 * 2 * kWinStride];
  __shared__ alignas(8) uint64_t bars[4 * 2];
  // cell inputs of the epilogue (U, U^n, rhs sum), staged by TMA with the windows: 2 x 4 KB per warp
  // Compiled in only with -DJXF_ROWS_TMA_IN=1: measured twice on the same box against per-lane loads issued at the top of
  // the iteration -- 7.55 vs 7.57 ms (profiles/r02i) and, with the lean halo path, 7.58 vs 7.47 ms (profiles/r02p): the
  // three extra TMA issues and their waits cost ~60 executed instructions per face-warp against the 15 LDG they replace
  constexpr bool kTmaIn = (EPI != 0) && (USE_TMA != 0) && (JXF_ROWS_TMA_IN != 0);
  __shared__ alignas(128) unsigned char in_raw[kTmaIn ? 4 * 2 * kInStride : 16];
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const long long gwarp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long ngwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long nrows = (long long)g.n1 * g.n2;
  const int G = ra.group_rows;
  const long long ngroups = (nrows + G - 1) / G;
  const int ipr = ra.iters_per_row;
  const int nA = g.nA;
  const bool tma_in = kTmaIn && ra.tma_in;
  const uint32_t in_s = smem_u32(in_raw) + (kTmaIn ? wid * 2 * kInStride : 0);
  const double step = (EPI ? (*a.dt) * a.dt_mult : 0.0);
  unsigned char* const win0 = win_raw + wid * 2 * kWinStride;      // this warp's two window buffers
  uint64_t* const bar0 = &bars[wid * 2];
  const uint32_t bar_s = smem_u32(bar0), win_s = smem_u32(win0);
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
  const int prev_lane = (lane + 31) & 31;
#if !defined(JXF_REFERENCE_ORDER) && JXF_ROWS_LANE_CARRY
  // lane carry of the cell-centred WENO weights: the option-free tuned instantiations only
  constexpr bool kLaneCarry = (RIEMANN == RIEMANN_HLLC_PLAIN) && ((RECON >> 1) != STENCIL_GENERIC);
  // lane 0's carries (the previous iteration's lane-31 weights) live in shared memory, not in registers of all lanes
  __shared__ double gcarry_s[4][16];
  double* const gcs = gcarry_s[wid];
#endif

  // stage the window of (row (i1n, i2n), iteration itn) into buffer b; `on` = false posts nothing (past the end)
  auto issue = [&](int b, int itn, int i1n, int i2n, bool on) {
    if (USE_TMA) {
      // TMA dims: (contiguous sweep axis, faster transverse (role 2), slower transverse (role 1), variable);
      // cells past the end of the row are zero-filled by the TMA unit.  One elected lane, predicated (no branch).
      const int c0 = ra.cA_off + 32 * itn - ra.shift, c1 = ra.c2_off + i2n, c2 = ra.c1_off + i1n;
      const uint32_t bar = bar_s + 8u * b, dst = win_s + (uint32_t)kWinStride * b;
      const int pred = (int)(on && lane == 0);
      const int tx = kWinBytes + (tma_in ? kInUBytes + kInRBytes + (EpiFlags<EPI>::blend(a) ? kInUBytes : 0) : 0);
      asm volatile(
          "{\n"
          ".reg .pred p;\n"
          "setp.ne.s32 p, %7, 0;\n"
          "@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %8;\n"
          "@p cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%2, {%3, %4, %5, %6}], [%1];\n"
          "}\n" ::"r"(dst), "r"(bar), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(0),
          "r"(pred), "r"(tx)
          : "memory");
      if (kTmaIn) {
        if (tma_in) {      // uniform
          const uint32_t ib = in_s + (uint32_t)kInStride * b;
          const int cu = ra.cA_off + 32 * itn - ra.lead;
          const int pun = pred && EpiFlags<EPI>::blend(a);
          asm volatile(
              "{\n"
              ".reg .pred p, q;\n"
              "setp.ne.s32 p, %9, 0;\n"
              "setp.ne.s32 q, %10, 0;\n"
              "@p cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%3, {%6, %7, %8, %11}], [%2];\n"
              "@q cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%1], [%4, {%6, %7, %8, %11}], [%2];\n"
              "}\n" ::"r"(ib), "r"(ib + (uint32_t)kInUStride), "r"(bar), "l"(reinterpret_cast<uint64_t>(&im.u)),
              "l"(reinterpret_cast<uint64_t>(&im.un)), "r"(0), "r"(cu), "r"(c1), "r"(c2), "r"(pred), "r"(pun), "r"(0)
              : "memory");
          // rhs accumulator: interior-only (or slab-local) coordinates
          asm volatile(
              "{\n"
              ".reg .pred p;\n"
              "setp.ne.s32 p, %6, 0;\n"
              "@p cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%2, {%3, %4, %5, %7}], [%1];\n"
              "}\n" ::"r"(ib + 2u * (uint32_t)kInUStride), "r"(bar), "l"(reinterpret_cast<uint64_t>(&im.rhs)), "r"(32 * itn),
              "r"(i2n), "r"(i1n), "r"(pred), "r"(0)
              : "memory");
        }
      }
    } else {
      if (on) {
        double* const wb = reinterpret_cast<double*>(win0 + b * kWinStride);
        const double* src = a.prims + i1n * g.s1 + i2n * g.s2 + (long long)(32 * itn - ra.shift) * g.sA;
        const int cmax = nA + 2 - (32 * itn - ra.shift);  // slots holding cells <= nA+2 are valid
#pragma unroll
        for (int v = 0; v < 5; ++v) {
          if (lane <= cmax) cp_async8(wb + v * kWinSlots + lane, src + v * g.vst + lane);
          if (lane < 6 && 32 + lane <= cmax) cp_async8(wb + v * kWinSlots + 32 + lane, src + v * g.vst + 32 + lane);
        }
      }
      cp_async_commit();
    }
  };

  for (long long group = gwarp; group < ngroups; group += ngwarps) {
    const long long row0 = group * G;
    const int nr = (int)min((long long)G, nrows - row0);
    const int i1_0 = (int)(row0 / g.n2);
    const int i2_0 = (int)(row0 - (long long)i1_0 * g.n2);
    int b = 0;
    bool next_landed = false;
    issue(0, 0, i1_0, i2_0, true);
    // ---- the row-opening faces f = 0 of this group, one row per lane (direct strided loads; the flux
    // code is called out of line here so the hot loop below holds the only inlined copy) -----------
    double F0[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    if (lane < nr) {
      const long long row = row0 + lane;
      const int k1 = (int)(row / g.n2);
      const int k2 = (int)(row - (long long)k1 * g.n2);
      face_flux_from_global<A, RECON, RIEMANN>(a.prims + k1 * g.s1 + k2 * g.s2 - 3 * g.sA, g.vst, g.sA, a.gamma, F0, a.limiter, a.fl);
    }
    int i1 = i1_0, i2 = i2_0;
    for (int r = 0; r < nr; ++r) {
      // next row (for the staging of its first window during this row's last iteration)
      int i1x = i1, i2x = i2 + 1;
      if (i2x == g.n2) {
        i2x = 0;
        ++i1x;
      }
      const bool more_rows = r + 1 < nr;
      const long long col_h = i1 * g.s1 + i2 * g.s2;
      const long long col_r = i1 * g.r1 + i2 * g.r2;
      double carry[5];
#pragma unroll
      for (int v = 0; v < 5; ++v) carry[v] = __shfl_sync(0xffffffffu, F0[v], r);
      for (int it = 0; it < ipr; ++it) {
        // stage the next iteration's window into the other buffer
        const bool last_it = it + 1 == ipr;
        issue(b ^ 1, last_it ? 0 : it + 1, last_it ? i1x : i1, last_it ? i2x : i2, !last_it || more_rows);
        const int f = 1 + 32 * it + lane;
        const bool act = f <= nA;
        const long long hidx = col_h + (long long)(f - 1) * g.sA;
        const long long ridx = col_r + (long long)(f - 1) * g.rA;
        CellIn<EPI> in;
        if (act && !tma_in) load_cell_in<EPI>(g, a, hidx, ridx, in);
        // wait for this iteration's window (in-place stages already waited for it before the previous stores)
        const double* const wb = reinterpret_cast<const double*>(win0 + b * kWinStride);
        if (!next_landed) {
          if (USE_TMA) {
            mbar_wait(bar0 + b, (phase_bits >> b) & 1u);
            phase_bits ^= (1u << b);
          } else {
            cp_async_wait<1>();
            __syncwarp();
          }
        }
        double F[5];
        {
          double w[5][6];
          const double* wl = wb + (ra.shift - 2) + lane;
#pragma unroll
          for (int v = 0; v < 5; ++v)
#pragma unroll
            for (int k = 0; k < 6; ++k) w[v][k] = wl[v * kWinSlots + k];
          // (idle tail lanes see the zero-filled end of the window: their NaN result is never used or stored)
#if !defined(JXF_REFERENCE_ORDER) && JXF_ROWS_LANE_CARRY
          if constexpr (kLaneCarry) {
            // cell-centred WENO weights of the as-is fields: this lane evaluates those of its face's RIGHT cell; the LEFT
            // cell's are the right-cell weights of the lane before (rotate-shuffle; lane 0: lane 31's of the previous
            // iteration, or -- first iteration of a row -- its own evaluation)
            ReconCarry<RECON> gr, gl;
            recon_g_right<A, RECON>(w, gr);
#pragma unroll
            for (int q = 0; q < ReconCarry<RECON>::N; ++q) {
              const double r0 = __shfl_sync(0xffffffffu, gr.g[q].g0, prev_lane);
              const double r1 = __shfl_sync(0xffffffffu, gr.g[q].g1, prev_lane);
              const double r2 = __shfl_sync(0xffffffffu, gr.g[q].g2, prev_lane);
              gl.g[q].g0 = r0; gl.g[q].g1 = r1; gl.g[q].g2 = r2;
              if (lane == 0) {           // take the carry, leave lane 31's weights of this iteration as the next one
                gl.g[q].g0 = gcs[3 * q]; gl.g[q].g1 = gcs[3 * q + 1]; gl.g[q].g2 = gcs[3 * q + 2];
                gcs[3 * q] = r0; gcs[3 * q + 1] = r1; gcs[3 * q + 2] = r2;
              }
            }
            if (it == 0 && lane == 0) recon_carry_init<A, RECON>(w, gl);      // the row's first cell: nobody's right cell
            face_flux_given<A, RECON>(w, a.gamma, F, gl, gr);
          } else
#endif
          face_flux<A, RECON, RIEMANN>(w, a.gamma, F, a.limiter, a.fl);
        }
        double rr[5];
#pragma unroll
        for (int v = 0; v < 5; ++v) {
          const double rot = __shfl_sync(0xffffffffu, F[v], prev_lane);   // lane 0 receives lane 31's flux
          rr[v] = ((lane == 0) ? carry[v] : rot) - F[v];
          carry[v] = rot;                                                // meaningful on lane 0: the next carry
        }
        if (kTmaIn) {
          if (tma_in) {      // the cell's inputs from this iteration's staged boxes (landed with the window)
            const double* ib = reinterpret_cast<const double*>(in_raw + (wid * 2 + b) * kInStride);
#pragma unroll
            for (int v = 0; v < 5; ++v) {
              in.U[v] = ib[v * kInCells + ra.lead + lane];
              in.rhs[v] = ib[2 * (kInUStride / 8) + v * 32 + lane];
            }
            if (EpiFlags<EPI>::blend(a)) {
#pragma unroll
              for (int v = 0; v < 5; ++v) in.Un[v] = ib[kInUStride / 8 + v * kInCells + ra.lead + lane];
            }
          }
        }
        if (EPI && a.inplace) {
          // prims_out aliases prims: the window staged for the next iteration overlaps the cells stored below (and
          // the next row's is read from global while this one stores) -- it must have landed first
          if (USE_TMA) {
            if (!last_it || more_rows) {
              mbar_wait(bar0 + (b ^ 1), (phase_bits >> (b ^ 1)) & 1u);
              phase_bits ^= (1u << (b ^ 1));
            }
          } else {
            cp_async_wait<0>();
            __syncwarp();
          }
          next_landed = true;
        }
        if (act) finalize_cell<EPI>(g, a, hidx, ridx, in, rr, step, red, f - 1, i1, i2);
        __syncwarp();          // all lanes are done with win[b] before it is refilled two iterations later
        b ^= 1;
      }
      i1 = i1x;
      i2 = i2x;
    }
    next_landed = false;
    if (!USE_TMA) cp_async_wait<0>();
  }
  if (EPI) {
    if (EpiFlags<EPI>::reduce(a)) red_commit(red, a.red);
  }
}
#else
template <int A, int RECON, int RIEMANN, int EPI, int USE_TMA>
__global__ void __launch_bounds__(128, JXF_MIN_BLOCKS)
sweep_rows(const __grid_constant__ SweepGeom g, const __grid_constant__ SweepArgs a, const RowsArgs ra, const __grid_constant__ CUtensorMap tmap,
           const __grid_constant__ RowsInMaps) {
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
    if (EpiFlags<EPI>::reduce(a)) red_commit(red, a.red);
  }
}

#endif  // JXF_ROWS_V1

}  // namespace jxf
