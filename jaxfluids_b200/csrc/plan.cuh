// plan.cuh -- host side shared by the translation units: the solver handle, launch accounting and the launch plan of
// one sweep (launch_sweep<A, RECON, RIEMANN, EPI>, explicitly instantiated in sweep_inst.cu).
#pragma once
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <cmath>
#include <new>

#include "sweep_kernels.cuh"

using namespace jxf;

int jxf_fail(int code, const char* fmt, ...);      // sets jxf_last_error(); defined in jxf_b200.cu
#define fail jxf_fail

inline int check_launch(const char* what) {
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
  double* peer_prims[6];   // jxf_set_peer_halo: the neighbours' output buffers of the NEXT stage call (peer-mapped), or null
  double* peer_cons[6];
  FaceData face_data;  // jxf_set_face_data: device pointers owned by the caller
  int has_face_data;
  int rows_group;      // JXF_ROWS_G=<1..32>: rows per warp work item of the rows kernel (tuning; 0 = automatic)
  bool no_lane_defer;  // default true; JXF_LANE_DEFER=1: the faces of the rows kernel's own axis are filled by a separate launch (A/B)
  bool lean_images;    // JXF_LEAN_IMAGES=1: mirror / periodic images of the rows kernel's own axis written inline (A/B)
  bool no_tma_in;      // JXF_NO_TMA_IN=1: the rows kernel's epilogue loads its cell inputs per lane (A/B only)
  bool no_plain;       // JXF_NO_PLAIN=1: never use the RIEMANN_HLLC_PLAIN / compile-time-flag instantiations (A/B only)
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

// TMA descriptor of a halo'd field buffer for the rows kernel (nullptr: use the cp.async loader); jxf_b200.cu
const CUtensorMap* get_rows_map(jxf_solver* s, const double* base);
// halo_fill_kernel on the faces in face_mask only (jxf_b200.cu); prims / cons: buffer bases
int launch_halo_faces(const jxf_solver* s, double* prims, double* cons, int face_mask, cudaStream_t st);
bool encode_rows_input_map(const jxf_solver* s, CUtensorMap* out, const double* base, bool is_rhs, int rhs_planes,
                           long long rhs_vst);

inline void set_role_bcs(SweepGeom& sg, const SweepArgs& a) {
  sg.bcA_hi = a.bc[2 * sg.axA]; sg.bcA_lo = a.bc[2 * sg.axA + 1];
  sg.bc1_hi = a.bc[2 * sg.ax1]; sg.bc1_lo = a.bc[2 * sg.ax1 + 1];
  sg.bc2_hi = a.bc[2 * sg.ax2]; sg.bc2_lo = a.bc[2 * sg.ax2 + 1];
}

template <int A, int RECON, int RIEMANN, int EPI>
int launch_sweep(const jxf_solver* s, SweepArgs a, cudaStream_t st) {
  const Geom& g = s->g;
  const int resident = s->num_sms * JXF_MIN_BLOCKS;   // CTAs of 128 threads resident on the device
  // fold the interior origin into the base pointers
  const long long h0 = g.off[0] * g.st[0] + g.off[1] * g.st[1] + g.off[2] * g.st[2];
  a.prims += h0;
  if (a.cons_in) a.cons_in += h0;
  if (a.cons_n) a.cons_n += h0;
  if (a.cons_out) a.cons_out += h0;
  if (a.prims_out) a.prims_out += h0;
  for (int f = 0; f < 6; ++f) {
    if (a.peer_prims[f]) a.peer_prims[f] += h0;
    if (a.peer_cons[f]) a.peer_cons[f] += h0;
  }
  SweepGeom sg;
  sg.axA = A;
  sg.nA = g.n[A];
  sg.sA = g.st[A];
  sg.rA = g.rst[A];
  sg.vst = g.vst;
  sg.rvst = a.rvst_slab > 0 ? a.rvst_slab : g.rvst;
  sg.i1_base = 0;
  sg.nearA_lo = s->cfg.nh;
  sg.nearA_hi = g.n[A] - s->cfg.nh;
  sg.leanA_lo = sg.leanA_hi = 0;
  // slab launches: a y / z sweep over the x planes [sub_lo, sub_lo + sub_n) only (x is role 1 of both in 3-D); the
  // field pointers move to the slab's first plane, the rhs pointer is the slab-sized accumulator as passed
  const bool slab = (A != 0) && a.sub_n > 0;
  if (slab) {
    const long long off = (long long)a.sub_lo * g.st[0];
    a.prims += off;
    if (a.cons_in) a.cons_in += off;
    if (a.cons_n) a.cons_n += off;
    if (a.cons_out) a.cons_out += off;
    if (a.prims_out) a.prims_out += off;
    sg.i1_base = a.sub_lo;
  }
  const int T1 = (A == 0) ? 1 : 0;       // slower transverse axis
  const int T2 = (A == 2) ? 1 : 2;       // faster transverse axis
  if (A != s->lane_axis) {
    // lanes along the contiguous axis C; the other transverse axis O is the slow one
    const int C = s->lane_axis;
    const int O = 3 - A - C;
    sg.ax1 = O; sg.n1 = g.n[O]; sg.s1 = g.st[O]; sg.r1 = g.rst[O];
    sg.ax2 = C; sg.n2 = g.n[C]; sg.s2 = g.st[C]; sg.r2 = g.rst[C];
    sg.n1_full = sg.n1;
    if (slab) {
      if (O != 0) return fail(JXF_ERR_UNSUPPORTED, "slab launch: x is not the slow transverse axis of this sweep");
      sg.n1 = a.sub_n;
    }
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
    ProfScope prof(s, A + 3 * (EPI ? 1 : 0), st);
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
    sg.n1_full = sg.n1;
    if (slab) {
      if (T1 != 0) return fail(JXF_ERR_UNSUPPORTED, "slab launch: x is not the slow transverse axis of this sweep");
      sg.n1 = a.sub_n;
    }
    set_role_bcs(sg, a);
    const long long rows = (long long)sg.n1 * sg.n2;
    const int nf = g.n[A] + 1;
    const long long total = rows * nf;
#if JXF_ROWS_KERNEL
    // production form whenever groups of >= 4 rows give every resident warp several work items.  (Measured, round 2:
    // widening this rule -- groups down to 1 row, e.g. the rows kernel for the 1024 rows of 2-D 1024^2 -- is SLOWER than
    // the contiguous kernel below there: 0.126 vs 0.096 ms per sweep with one partial wave of warps, and G = 2 instead
    // of 4 at 256^3 costs 12 %: profiles/r02j_*.json against r02b_*.json.)
    const long long warps_resident = 4LL * resident;
    // (an in-place epilogue needs the staged windows of this kernel: forced whatever the slab's row count)
    if ((s->force_rows || (EPI && a.inplace) || rows / 4 >= warps_resident * 2) && g.n[A] >= 32) {
      RowsArgs ra;
      ra.iters_per_row = (g.n[A] + 31) / 32;
      // one group per warp, 4 warps per CTA, many more CTAs than resident slots: the hardware block
      // scheduler balances the tail (a static groups-per-warp split left ~8 % of the warps idle at the end)
      int G = 8;
      while (G > 4 && rows / G < warps_resident * 16) G >>= 1;
      if (s->rows_group > 0) G = s->rows_group;            // JXF_ROWS_G (tuning)
      ra.group_rows = G;
      ra.shift = ((g.off[A] - 2) & 1) ? 3 : 2;
      ra.cA_off = g.off[A];
      ra.c1_off = g.off[T1] + (slab ? a.sub_lo : 0);
      ra.c2_off = g.off[T2];
      ra.tma_dim1_is_role = 2;
      const long long groups = (rows + G - 1) / G;
      const long long blocks = std::min<long long>((groups + 3) / 4, 1LL << 30);
      const CUtensorMap* map = get_rows_map(const_cast<jxf_solver*>(s), a.prims - h0 - (slab ? (long long)a.sub_lo * g.st[0] : 0));
      // the epilogue's cell inputs (U, U^n, the earlier axes' rhs sum) staged by TMA next to the windows
      RowsInMaps im;
      memset(&im, 0, sizeof(im));
      ra.tma_in = 0;
      ra.lead = g.off[A] & 1;             // the U / U^n boxes start `lead` cells before the iteration's first cell (16 B)
      if (JXF_ROWS_TMA_IN && EPI && map && !s->no_tma_in && a.has_prev && a.rhs && a.cons_in) {
        const long long slab_off = slab ? (long long)a.sub_lo * g.st[0] : 0;
        bool ok = encode_rows_input_map(s, &im.u, a.cons_in - h0 - slab_off, false, 0, 0) &&
                  encode_rows_input_map(s, &im.rhs, a.rhs, true, slab ? a.sub_n : g.n[0], sg.rvst);
        if (ok && a.blend) ok = a.cons_n && encode_rows_input_map(s, &im.un, a.cons_n - h0 - slab_off, false, 0, 0);
        ra.tma_in = ok ? 1 : 0;
      }
      // The halo images of the two faces of THIS axis come from the row ends: one warp iteration in eight takes the
      // out-of-line boundary-cell path with 5 active lanes (0.31 ms of the 7.27 ms launch at 512^3).  JXF_LANE_DEFER=1
      // leaves them to one halo_fill launch on those two faces after the sweep -- measured SLOWER in total (the separate
      // launch takes 0.40 ms: 2.6 M row ends x 15 fields, every access a 40-byte segment in its own DRAM page, where the
      // fused stores land in the lines the row's last cells are being written to; profiles/r02u_ab_lane_defer.txt), so it
      // is off by default.  Never with peer-mapped stores on these faces, never for slab launches.
      // SYMMETRY / PERIODIC faces of this axis without boundary data or peer stores: images written inline by
      // finalize_cell (SweepGeom::leanA_*), the out-of-line path skips them
      if (EPI && a.fuse_halo && s->lean_images && !a.has_face_data) {
        const int khi = a.bc[2 * A], klo = a.bc[2 * A + 1];
        const bool per = khi == JXF_BC_PERIODIC && klo == JXF_BC_PERIODIC;
        if (per || khi == JXF_BC_SYMMETRY) { sg.leanA_hi = per ? 2 : 1; sg.bcA_hi = JXF_BC_INACTIVE; sg.nearA_hi = g.n[A]; }
        if (per || klo == JXF_BC_SYMMETRY) { sg.leanA_lo = per ? 2 : 1; sg.bcA_lo = JXF_BC_INACTIVE; sg.nearA_lo = 0; }
      }
      int defer_mask = 0;
      if (EPI && a.fuse_halo && !slab && !s->no_lane_defer && !a.peer_prims[2 * A] && !a.peer_prims[2 * A + 1]) {
        for (int f = 2 * A; f < 2 * A + 2; ++f)
          if (a.bc[f] != JXF_BC_INACTIVE && a.bc[f] != JXF_BC_NEIGHBOR) defer_mask |= 1 << f;
        if (defer_mask) {
          sg.bcA_hi = sg.bcA_lo = JXF_BC_INACTIVE;
          sg.nearA_lo = 0;
          sg.nearA_hi = g.n[A];
        }
      }
      {
        ProfScope prof(s, A + 3 * (EPI ? 1 : 0), st);
        if (map) {
          sweep_rows<A, RECON, RIEMANN, EPI, 1><<<(unsigned)blocks, 128, 0, st>>>(sg, a, ra, *map, im);
        } else {
          CUtensorMap dummy;
          memset(&dummy, 0, sizeof(dummy));
          sweep_rows<A, RECON, RIEMANN, EPI, 0><<<(unsigned)blocks, 128, 0, st>>>(sg, a, ra, dummy, im);
        }
      }
      if (int rc = check_launch("sweep_rows")) return rc;
      if (defer_mask) return launch_halo_faces(s, a.prims_out - h0, a.cons_out - h0, defer_mask, st);
      return JXF_OK;
    }
#endif
    if (EPI && a.inplace)
      return fail(JXF_ERR_UNSUPPORTED, "in-place stage: the last sweep must run in the rows kernel (grid too small)");
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
    ProfScope prof(s, A + 3 * (EPI ? 1 : 0), st);
    sweep_contig<A, RECON, RIEMANN, EPI><<<(unsigned)std::max<long long>(1, blocks), 128, 0, st>>>(sg, a, total);
  }
  return check_launch("sweep");
}

// The instantiations of launch_sweep: X(A, RECON, RIEMANN, EPI).  Every (A, RECON) pair has the four (RIEMANN, EPI)
// combinations with run-time options; the tuned stencils (RECON 0..3) add the option-free HLLC + EINFELDT set with the
// run-time-flag epilogue (1) and the four compile-time-flag epilogues (2..5).
#define JXF_SWEEPS_OF(X, A, R) X(A, R, 0, 0) X(A, R, 0, 1) X(A, R, 1, 0) X(A, R, 1, 1)
#define JXF_SWEEPS_PLAIN_OF(X, A, R) X(A, R, 2, 0) X(A, R, 2, 1) X(A, R, 2, 2) X(A, R, 2, 3) X(A, R, 2, 4) X(A, R, 2, 5)
#define JXF_SWEEPS_TUNE(X, A) X(A, 1, 0, 0) X(A, 1, 0, 1) JXF_SWEEPS_PLAIN_OF(X, A, 1)
