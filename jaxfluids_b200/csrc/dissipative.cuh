// dissipative.cuh -- viscous + heat flux of the single-phase path (included by jxf_b200.cu).
//
// ref (file:line under /root/reference/src/jaxfluids/):
//   solvers/source_term_solver.py:188-250 (heat flux), :258-345 (viscous flux), :405-470 (velocity
//   gradient at faces), :503-533 (tau), :535-582 (d/dx_i at the faces of axis j);
//   stencils/derivative/deriv_face_4.py, deriv_center_4.py, stencils/reconstruction/central/central_4.py;
//   solvers/space_solver.py:567-599 (folded into the face flux before the divergence);
//   halos/outer/material.py:289-383 + boundary_condition.py:128-179, :607-655 (edge halos).
//
// Face flux of axis A between cells i and i+1, from the four cells i-1..i+2 along A:
//   d u_c / d x_A        : 1/dx (1/24 (u_{i-1} - u_{i+2}) + 27/24 (u_{i+1} - u_i))          (face derivative)
//   d u_c / d x_t, t != A: central_4 reconstruction along A of the CELL-CENTRE derivative
//                          1/dx_t (1/12 (u_{-2} - u_{+2}) + 8/12 (u_{+1} - u_{-1})) taken along t
//   tau_k = mu (du_A/dx_k + du_k/dx_A),  tau_A += (bulk - 2/3 mu) div u,  u.tau with central_4 face velocities,
//   q = -lambda dT/dx_A (face derivative), T = p / (rho R).
// mu, bulk, lambda are constants here (transport model CUSTOM with float values, PRANDTL with a CUSTOM
// viscosity), so the reference's T-at-face reconstruction only feeds constants and is not evaluated.
//
// Two kernels, the same split as the convective sweeps:
//   visc_march<A>: A is NOT the contiguous axis.  Thread = one column, lanes along the contiguous axis,
//       marching along A with a rolling 4-cell window of {u, v, w, T, six transverse cell-centre
//       derivatives}; every cell's derivatives and every face flux are computed once.
//   visc_rows<A>:  A IS the contiguous axis.  CTA = one row segment, thread = one cell: each thread forms
//       its cell's data, the four-cell stencils and the flux differences go through shared memory.
// Both are streaming kernels bound by HBM: 40 B/cell of primitives in + 32 B/cell rhs in/out.
#pragma once

namespace jxf {

#ifdef JXF_REFERENCE_ORDER
__device__ __forceinline__ double visc_rcp(double a) { return 1.0 / a; }
#else
__device__ __forceinline__ double visc_rcp(double a) { return rcp_fast(a); }
#endif

struct ViscArgs {
  const double* prims;
  double* rhs;
  double mu1, mu2;          // mu, bulk - 2/3 mu
  double lambda;            // thermal conductivity
  double gas_constant;
  double inv_dxA, inv_dx1, inv_dx2;   // 1/dx of the three ROLE axes (A, 1, 2)
  int visc, heat, heat_prod;
  int active_mask;          // bit i = physical axis i active
  int accumulate;
  int chunk_len;            // march: cells per chunk along A
  int seg_len;              // rows: cells per CTA along A
};

struct VCell {
  double u[3];
  double T;
  double d1[3];             // d u_c / d x_(role 1) at the cell centre
  double d2[3];             // d u_c / d x_(role 2)
};

// raw operands of one cell: own (rho, u, v, w, p) and the +-1, +-2 neighbours of the velocities along the two
// transverse role axes.  Loading (vraw_load) and reducing (vcell_from_raw) are separate so that the marching
// kernel can post the loads of the NEXT cell before it works on the current face (software pipeline).
struct VRaw {
  double own[5];
  double n1[3][4];          // velocity k at -2, -1, +1, +2 along role 1
  double n2[3][4];
};

__device__ __forceinline__ void vraw_load(const SweepGeom& g, const ViscArgs& a, long long idx, VRaw& r) {
  const double* p = a.prims + idx;
  r.own[0] = r.own[4] = 1.0;
  if (a.heat) {
    r.own[0] = p[0];
    r.own[4] = p[4 * g.vst];
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) r.own[1 + k] = p[(1 + k) * g.vst];
  if (a.visc) {
    if (g.n1 > 1) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double* q = p + (1 + k) * g.vst;
        r.n1[k][0] = q[-2 * g.s1]; r.n1[k][1] = q[-g.s1]; r.n1[k][2] = q[g.s1]; r.n1[k][3] = q[2 * g.s1];
      }
    }
    if (g.n2 > 1) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double* q = p + (1 + k) * g.vst;
        r.n2[k][0] = q[-2 * g.s2]; r.n2[k][1] = q[-g.s2]; r.n2[k][2] = q[g.s2]; r.n2[k][3] = q[2 * g.s2];
      }
    }
  }
}

__device__ __forceinline__ void vcell_from_raw(const SweepGeom& g, const ViscArgs& a, const VRaw& r, VCell& c) {
  constexpr double c0 = 1.0 / 12.0, c1 = 8.0 / 12.0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    c.u[k] = r.own[1 + k];
    c.d1[k] = 0.0;
    c.d2[k] = 0.0;
  }
  c.T = a.heat ? r.own[4] * visc_rcp(r.own[0] * a.gas_constant) : 0.0;
  if (a.visc) {
    if (g.n1 > 1) {
#pragma unroll
      for (int k = 0; k < 3; ++k) c.d1[k] = a.inv_dx1 * fma(c1, r.n1[k][2] - r.n1[k][1], c0 * (r.n1[k][0] - r.n1[k][3]));
    }
    if (g.n2 > 1) {
#pragma unroll
      for (int k = 0; k < 3; ++k) c.d2[k] = a.inv_dx2 * fma(c1, r.n2[k][2] - r.n2[k][1], c0 * (r.n2[k][0] - r.n2[k][3]));
    }
  }
}

// cell-centre data of the cell at element offset `idx` (relative to the interior origin)
__device__ __forceinline__ void load_vcell(const SweepGeom& g, const ViscArgs& a, long long idx, VCell& c) {
  VRaw r;
  vraw_load(g, a, idx, r);
  vcell_from_raw(g, a, r, c);
}

__device__ __forceinline__ double central4(double a, double b, double c, double d) {
  return fma(9.0 / 16.0, b + c, (-1.0 / 16.0) * (a + d));
}
__device__ __forceinline__ double dface4(double a, double b, double c, double d) {
  return fma(27.0 / 24.0, c - b, (1.0 / 24.0) * (a - d));
}

// dissipative part of the face flux, components (momentum x, y, z, energy), sign as it enters the
// convective flux: Fd = (-tau, -u.tau + q)
template <int A>
__device__ __forceinline__ void dissipative_face_flux(const SweepGeom& g, const ViscArgs& a, const VCell& c0,
                                                      const VCell& c1, const VCell& c2, const VCell& c3,
                                                      double (&F)[4]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) F[k] = 0.0;
  if (a.visc) {
    // vg[c][i] = d u_c / d x_i at the face; i indexes PHYSICAL axes
    double vg[3][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
      for (int i = 0; i < 3; ++i) vg[k][i] = 0.0;
      vg[k][A] = a.inv_dxA * dface4(c0.u[k], c1.u[k], c2.u[k], c3.u[k]);
      const double t1 = central4(c0.d1[k], c1.d1[k], c2.d1[k], c3.d1[k]);
      const double t2 = central4(c0.d2[k], c1.d2[k], c2.d2[k], c3.d2[k]);
      // roles -> physical axes are compile-time for a given A and lane axis layout, but the role table
      // is in g: ax1/ax2 are uniform, so these selects are cheap
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        if (i != A) vg[k][i] = (i == g.ax1) ? t1 : ((i == g.ax2) ? t2 : 0.0);
      }
    }
    const bool act[3] = {(a.active_mask & 1) != 0, (a.active_mask & 2) != 0, (a.active_mask & 4) != 0};
    double tau[3];
    double div = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      tau[k] = act[k] ? a.mu1 * (vg[A][k] + vg[k][A]) : 0.0;
      if (act[k]) div += vg[k][k];
    }
    tau[A] = fma(a.mu2, div, tau[A]);
    double vt = 0.0;
    if (a.heat_prod) {
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (act[k]) vt = fma(tau[k], central4(c0.u[k], c1.u[k], c2.u[k], c3.u[k]), vt);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) F[k] = -tau[k];
    F[3] = -vt;
  }
  if (a.heat) F[3] = fma(-a.lambda, a.inv_dxA * dface4(c0.T, c1.T, c2.T, c3.T), F[3]);
}

__device__ __forceinline__ void dissipative_update(const SweepGeom& g, const ViscArgs& a, long long ridx,
                                                   const double (&Flo)[4], const double (&Fhi)[4]) {
  if (!a.accumulate) a.rhs[ridx] = 0.0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    double* r = a.rhs + ridx + (1 + k) * g.rvst;
    const double d = Flo[k] - Fhi[k];
    *r = a.accumulate ? fma(a.inv_dxA, d, *r) : a.inv_dxA * d;
  }
}

// CTA = 32 lanes along the contiguous axis x 4 columns along the other transverse axis: the +-1, +-2
// neighbours a cell-centre derivative reads along that axis are mostly the CTA's own columns (L1 hits)
template <int A>
__global__ void __launch_bounds__(128, 3) visc_march(const __grid_constant__ SweepGeom g, const __grid_constant__ ViscArgs a) {
  const int i2 = blockIdx.x * 32 + threadIdx.x;
  const int i1 = blockIdx.y * 4 + threadIdx.y;
  if (i1 >= g.n1 || i2 >= g.n2) return;
  const int f0 = blockIdx.z * a.chunk_len;
  const int f1 = min(f0 + a.chunk_len, g.nA);
  const long long col_h = i1 * g.s1 + i2 * g.s2;
  const long long col_r = i1 * g.r1 + i2 * g.r2;
  VCell c0, c1, c2, c3;
  load_vcell(g, a, col_h + (long long)(f0 - 2) * g.sA, c0);
  load_vcell(g, a, col_h + (long long)(f0 - 1) * g.sA, c1);
  load_vcell(g, a, col_h + (long long)f0 * g.sA, c2);
  VRaw nx;
  vraw_load(g, a, col_h + (long long)(f0 + 1) * g.sA, nx);        // cell f0+1, consumed by the first iteration
  double Fp[4] = {0.0, 0.0, 0.0, 0.0};
  for (int f = f0; f <= f1; ++f) {
    vcell_from_raw(g, a, nx, c3);                                  // cell f+1 <= n+1 < n+nh
    // post the next cell's loads and this cell's rhs loads now; they are consumed one face of arithmetic later
    if (f < f1) vraw_load(g, a, col_h + (long long)(f + 2) * g.sA, nx);
    const long long ridx = col_r + (long long)(f - 1) * g.rA;
    double r_old[4] = {0.0, 0.0, 0.0, 0.0};
    if (f > f0 && a.accumulate) {
#pragma unroll
      for (int k = 0; k < 4; ++k) r_old[k] = a.rhs[ridx + (1 + k) * g.rvst];
    }
    double F[4];
    dissipative_face_flux<A>(g, a, c0, c1, c2, c3, F);
    if (f > f0) {
      if (!a.accumulate) a.rhs[ridx] = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) a.rhs[ridx + (1 + k) * g.rvst] = fma(a.inv_dxA, Fp[k] - F[k], r_old[k]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) Fp[k] = F[k];
    c0 = c1;
    c1 = c2;
    c2 = c3;
  }
}

// shared layout per row of the CTA: 10 cell quantities + 4 flux components, each blockDim.x doubles.
// CTA = blockDim.y consecutive rows x one segment of the contiguous axis (neighbouring rows share the lines
// their transverse derivatives read)
#ifndef JXF_VISC_ROWS_MINB
#define JXF_VISC_ROWS_MINB 1
#endif
template <int A>
__global__ void __launch_bounds__(256, JXF_VISC_ROWS_MINB) visc_rows(const __grid_constant__ SweepGeom g, const __grid_constant__ ViscArgs a) {
  extern __shared__ double vs_all[];
  const int B = blockDim.x;
  const int tid = threadIdx.x;
  double* const vs = vs_all + (size_t)threadIdx.y * 14 * B;
  const long long nrows = (long long)g.n1 * g.n2;
  const long long row = blockIdx.x * (long long)blockDim.y + threadIdx.y;
  const bool live = row < nrows;
  const int i1 = live ? (int)(row / g.n2) : 0;
  const int i2 = live ? (int)(row - (long long)i1 * g.n2) : 0;
  const int s0 = blockIdx.y * a.seg_len;                 // first cell of this segment
  const int L = min(a.seg_len, g.nA - s0);               // cells in this segment
  const long long col_h = i1 * g.s1 + i2 * g.s2;
  const long long col_r = i1 * g.r1 + i2 * g.r2;
  const int k = s0 - 2 + tid;                            // this thread's cell
  const bool owns = live && tid >= 2 && tid <= L + 1;    // cell k is updated by this thread
  const long long ridx = col_r + (long long)k * g.rA;
  double r_old[4] = {0.0, 0.0, 0.0, 0.0};
  if (owns && a.accumulate) {                            // posted with the stencil loads, consumed after both barriers
#pragma unroll
    for (int q = 0; q < 4; ++q) r_old[q] = a.rhs[ridx + (1 + q) * g.rvst];
  }
  if (live && tid < L + 4) {
    VCell c;
    load_vcell(g, a, col_h + (long long)k * g.sA, c);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      vs[q * B + tid] = c.u[q];
      vs[(4 + q) * B + tid] = c.d1[q];
      vs[(7 + q) * B + tid] = c.d2[q];
    }
    vs[3 * B + tid] = c.T;
  }
  __syncthreads();
  double* fl = vs + 10 * B;
  if (live && tid >= 1 && tid <= L + 1) {                // face between cells k and k+1
    VCell c[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = tid - 1 + j;
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        c[j].u[q] = vs[q * B + t];
        c[j].d1[q] = vs[(4 + q) * B + t];
        c[j].d2[q] = vs[(7 + q) * B + t];
      }
      c[j].T = vs[3 * B + t];
    }
    double F[4];
    dissipative_face_flux<A>(g, a, c[0], c[1], c[2], c[3], F);
#pragma unroll
    for (int q = 0; q < 4; ++q) fl[q * B + tid] = F[q];
  }
  __syncthreads();
  if (owns) {                                            // faces computed by tid-1 (low) and tid (high)
    if (!a.accumulate) a.rhs[ridx] = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      a.rhs[ridx + (1 + q) * g.rvst] = fma(a.inv_dxA, fl[q * B + tid - 1] - fl[q * B + tid], r_old[q]);
  }
}

// ---------------------------------------------------------------------------
// edge halos (halos/outer/material.py:289-383)
// ---------------------------------------------------------------------------
struct EdgeArgs {
  double* prims;
  double* cons;
  double gamma;
  int bc[6];
};

// faces of the 12 edges in the reference's order (domain/__init__.py:9-15), as face ids
// (east 0, west 1, north 2, south 3, top 4, bottom 5)
__constant__ int kEdgeFaces[12][2] = {{1, 3}, {1, 2}, {0, 2}, {0, 3}, {3, 5}, {2, 5},
                                      {3, 4}, {2, 4}, {0, 5}, {1, 5}, {0, 4}, {1, 4}};

__device__ __forceinline__ int edge_src_index(int kind, bool hi, int l, int nh, int ext) {
  // buffer index along one axis of the edge for halo layer l (in increasing buffer index):
  //   kind 0: the halo itself; 1: adjacent interior layers, same order ("_1" slices, halo_slices.py:98-147);
  //   2: PERIODIC, the interior layers next to the OPPOSITE face; 3: SYMMETRY, adjacent layers mirrored
  switch (kind) {
    case 0: return hi ? ext - nh + l : l;
    case 1: return hi ? ext - 2 * nh + l : nh + l;
    case 2: return hi ? nh + l : ext - 2 * nh + l;
    default: return hi ? ext - nh - 1 - l : 2 * nh - 1 - l;
  }
}

__global__ void __launch_bounds__(128) halo_fill_edges_kernel(const Geom g, const EdgeArgs a) {
  const int e = blockIdx.y;
  const int fa = kEdgeFaces[e][0], fb = kEdgeFaces[e][1];
  const int axa = fa >> 1, axb = fb >> 1;
  if (g.n[axa] <= 1 || g.n[axb] <= 1) return;
  const int ta = a.bc[fa], tb = a.bc[fb];
  if (ta == JXF_BC_NEIGHBOR || tb == JXF_BC_NEIGHBOR) return;      // inter-block edges: not handled here
  const int axr = 3 - axa - axb;                                     // running axis
  const int nr = g.n[axr], nh = g.nh;
  const bool hia = (fa & 1) == 0, hib = (fb & 1) == 0;
  // boundary_condition.py:128-179: the first PERIODIC / SYMMETRY face decides
  int which = -1;   // 0: face a decides, 1: face b decides, -1: ANY_ANY
  if (ta == JXF_BC_PERIODIC || ta == JXF_BC_SYMMETRY) which = 0;
  else if (tb == JXF_BC_PERIODIC || tb == JXF_BC_SYMMETRY) which = 1;
  const int tdec = which == 0 ? ta : tb;
  const long long total = (long long)nh * nh * nr;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    const int ir = (int)(q % nr);
    const long long q1 = q / nr;
    const int lb = (int)(q1 % nh);
    const int la = (int)(q1 / nh);
    const long long run = (long long)(ir + g.off[axr]) * g.st[axr];
    const long long dst = run + (long long)edge_src_index(0, hia, la, nh, g.ext[axa]) * g.st[axa] +
                          (long long)edge_src_index(0, hib, lb, nh, g.ext[axb]) * g.st[axb];
    double p[5];
    if (which < 0) {
      const long long s1 = run + (long long)edge_src_index(1, hia, la, nh, g.ext[axa]) * g.st[axa] +
                           (long long)edge_src_index(0, hib, lb, nh, g.ext[axb]) * g.st[axb];
      const long long s2 = run + (long long)edge_src_index(0, hia, la, nh, g.ext[axa]) * g.st[axa] +
                           (long long)edge_src_index(1, hib, lb, nh, g.ext[axb]) * g.st[axb];
#pragma unroll
      for (int v = 0; v < 5; ++v) p[v] = 0.5 * (a.prims[s1 + v * g.vst] + a.prims[s2 + v * g.vst]);
    } else {
      const int kind = (tdec == JXF_BC_PERIODIC) ? 2 : 3;
      const int ka = which == 0 ? kind : 0, kb = which == 1 ? kind : 0;
      const long long s = run + (long long)edge_src_index(ka, hia, la, nh, g.ext[axa]) * g.st[axa] +
                          (long long)edge_src_index(kb, hib, lb, nh, g.ext[axb]) * g.st[axb];
#pragma unroll
      for (int v = 0; v < 5; ++v) p[v] = a.prims[s + v * g.vst];
      if (tdec == JXF_BC_SYMMETRY) {
        const int ax = which == 0 ? axa : axb;
        p[1 + ax] = p[1 + ax] * -1.0;
      }
    }
    double c[5];
    cons_from_prims(p, a.gamma, c);
#pragma unroll
    for (int v = 0; v < 5; ++v) {
      a.prims[dst + v * g.vst] = p[v];
      if (a.cons) a.cons[dst + v * g.vst] = c[v];
    }
  }
}

__global__ void __launch_bounds__(256) temperature_kernel(const double* __restrict__ prims, double* __restrict__ T,
                                                          long long vst, double gas_constant) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < vst; i += (long long)gridDim.x * blockDim.x)
    T[i] = prims[i + 4 * vst] / (prims[i] * gas_constant);
}

}  // namespace jxf
