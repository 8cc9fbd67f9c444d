/* jxf_b200.h -- C ABI of the B200-native (sm_100a) convective-RHS / SSP-RK path
 * that stands in for JAX-Fluids' single-phase space solver.
 *
 * Every entry point is enqueue-only on the caller's stream (`stream` is a
 * cudaStream_t passed as void*), never allocates, frees or retains caller
 * memory, and returns 0 on success or a negative jxf_status; the message of the
 * last failure on the calling thread is available from jxf_last_error().
 * All device buffers are fp64, C-order (5, X, Y, Z) with `nh` halo cells on
 * both sides of every active axis and extent 1 on inactive axes -- the
 * reference's buffer layout (initialization/helper_functions.py:49).
 * The rhs buffer is interior-only, (5, Nx, Ny, Nz) (space_solver.py:489).
 *
 * Citations "ref:" are file:line under /root/reference/src/jaxfluids/.
 * The XLA-FFI binding a maintainer would add on the reference side is shown in
 * INTEGRATION.md; it forwards to exactly these symbols.
 */
#ifndef JXF_B200_H
#define JXF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum jxf_status {
  JXF_OK = 0,
  JXF_ERR_BAD_ARG = -1,
  JXF_ERR_UNSUPPORTED = -2,   /* valid reference option that this path does not implement */
  JXF_ERR_CUDA = -3
} jxf_status;

/* ref: stencils/__init__.py:15-19 + godunov.reconstruction_variable (read_conservatives.py:126-203) */
enum { JXF_RECON_PRIMITIVE = 0, JXF_RECON_CHAR_PRIMITIVE = 1,
       /* high_order_godunov.py:282-296, :404-417: run in the generic (reference-order) kernel instantiations */
       JXF_RECON_CONSERVATIVE = 2, JXF_RECON_CHAR_CONSERVATIVE = 3 };
/* ref: godunov/frozen_state, flux_splitting/frozen_state (solvers/__init__.py:5-7; eigendecomposition.py:146-276);
 * ROE runs in the generic kernel instantiations */
enum { JXF_FROZEN_ARITHMETIC = 0, JXF_FROZEN_ROE = 1 };
/* ref: stencils/reconstruction/shock_capturing/weno/weno5_z.py, weno5_js.py (DICT_SPATIAL_RECONSTRUCTION): the two
 * tuned forms.  Ids >= 2: the other stencils of the reference that fit the kernels' 6-cell window, evaluated in
 * the reference's operation order by one generic set of kernel instantiations (weno/weno1_js.py, weno3_js.py,
 * weno3_z.py, weno3_n.py, teno/teno5.py, teno/teno6.py, weno/weno6_cu.py, muscl/muscl3.py with stencils/limiter.py,
 * reconstruction/central/central_2.py). */
enum { JXF_STENCIL_WENO5Z = 0, JXF_STENCIL_WENO5JS = 1, JXF_STENCIL_WENO1 = 2, JXF_STENCIL_WENO3JS = 3,
       JXF_STENCIL_WENO3Z = 4, JXF_STENCIL_TENO5 = 5, JXF_STENCIL_WENO6CU = 6, JXF_STENCIL_KOREN = 7, JXF_STENCIL_MC = 8,
       JXF_STENCIL_MINMOD = 9, JXF_STENCIL_SUPERBEE = 10, JXF_STENCIL_VANALBADA = 11, JXF_STENCIL_VANLEER = 12,
       JXF_STENCIL_WENO3N = 13, JXF_STENCIL_CENTRAL2 = 14, JXF_STENCIL_TENO6 = 15,
       JXF_STENCIL_TENO5A = 16 /* teno/teno5_a.py */, JXF_STENCIL_TENO6A = 17 /* teno/teno6_a.py */ };
/* ref: solvers/riemann_solvers/__init__.py:16-34 */
enum { JXF_RIEMANN_HLLC = 0, JXF_RIEMANN_RUSANOV = 1, JXF_RIEMANN_HLL = 2 /* HLL.py; uses jxf_config.signal_speed */,
       JXF_RIEMANN_HLLCLM = 3 /* HLLCLM.py (low-Mach HLLC, Fleischmann et al. 2020); uses jxf_config.signal_speed */,
       JXF_RIEMANN_AUSMP = 4 /* AUSMP.py (AUSM+) */ };
/* ref: solvers/riemann_solvers/signal_speeds.py (DICT_SIGNAL_SPEEDS); DAVIS2 is marked not working upstream */
enum { JXF_SIGNAL_EINFELDT = 0, JXF_SIGNAL_ARITHMETIC = 1, JXF_SIGNAL_RUSANOV = 2, JXF_SIGNAL_DAVIS = 3, JXF_SIGNAL_TORO = 4 };
/* ref: solvers/convective_fluxes/__init__.py:7-12; solvers/__init__.py:1-3 (TUPLE_FLUX_SPLITTING) */
enum { JXF_SOLVER_GODUNOV = 0, JXF_SOLVER_FLUX_SPLITTING = 1 };
enum { JXF_FS_ROE = 1, JXF_FS_CLLF = 2, JXF_FS_LLF = 3 };
/* ref: time_integration/__init__.py:6-11 */
enum { JXF_INT_EULER = 0, JXF_INT_RK2 = 1, JXF_INT_RK3 = 2, JXF_INT_RK2_LS4 = 3 /* RK2_LS4.py: 4 stages */ };
/* ref: halos/outer/__init__.py:1-7; NEIGHBOR = face owned by another rank (halos/inner/material.py:30-93) */
enum { JXF_BC_INACTIVE = 0, JXF_BC_PERIODIC = 1, JXF_BC_SYMMETRY = 2, JXF_BC_ZEROGRADIENT = 3, JXF_BC_NEIGHBOR = 4,
       JXF_BC_WALL = 5      /* ref: halos/outer/material.py:473-520, constant wall_velocity_callable */,
       JXF_BC_DIRICHLET = 6 /* ref: halos/outer/material.py:732-798, constant primitives_callable   */ };
/* face order of the reference: domain/__init__.py:5-7 */
enum { JXF_EAST = 0, JXF_WEST = 1, JXF_NORTH = 2, JXF_SOUTH = 3, JXF_TOP = 4, JXF_BOTTOM = 5 };

typedef struct jxf_config {
  int32_t n[3];          /* interior cells of this block per axis (1 = inactive axis)            */
  int32_t nh;            /* conservatives.halo_cells (>= 3, ref: weno5_base.py:18)               */
  double  inv_dx[3];     /* 1/dx per axis, as the reference forms it (domain_information.py:290) */
  double  dx_min;        /* min over active axes (domain_information.py:697-702)                 */
  double  gamma;         /* IdealGas specific_heat_ratio                                         */
  double  cfl;           /* time_integration.CFL                                                 */
  double  fixed_dt;      /* time_integration.fixed_timestep, 0 = CFL based                       */
  int32_t recon;         /* JXF_RECON_*     (the stencil is `stencil` below)                     */
  int32_t riemann;       /* JXF_RIEMANN_*                                                        */
  int32_t signal_speed;  /* JXF_SIGNAL_*                                                         */
  int32_t integrator;    /* JXF_INT_*                                                            */
  int32_t bc[6];         /* JXF_BC_* per face, order east,west,north,south,top,bottom            */
  /* dissipative fluxes (ref: active_physics is_viscous_flux / is_heat_flux / is_viscous_heat_production,
   * conservatives/dissipative_fluxes = CENTRAL4 x3, material_properties/transport; source_term_solver.py:188-345) */
  int32_t viscous_flux;             /* 0/1                                                        */
  int32_t heat_flux;                /* 0/1                                                        */
  int32_t viscous_heat_production;  /* 0/1: u.tau in the energy flux (default 1)                  */
  int32_t stencil;                  /* JXF_STENCIL_*: godunov.reconstruction_stencil (0 = WENO5-Z)    */
  double  dynamic_viscosity;        /* transport/dynamic_viscosity, model CUSTOM (constant)       */
  double  bulk_viscosity;           /* transport/bulk_viscosity                                   */
  double  thermal_conductivity;     /* constant lambda: CUSTOM value, or cp*mu/Pr for PRANDTL     */
  double  gas_constant;             /* equation_of_state/specific_gas_constant (T = p/(rho R))    */
  /* ref: conservatives/positivity (solvers/positivity/limiter_interpolation.py:77-209, SINGLE-PHASE branch):
   * reconstructed states with density < 1e-12 or pressure < 1e-10 fall back to the first-order state */
  int32_t interpolation_limiter;    /* 0/1: positivity/is_interpolation_limiter                   */
  int32_t limit_velocity;           /* 0/1: positivity/limit_velocity (all primitives, not just rho, p) */
  double  wall_velocity[6][3];      /* (u, v, w) of the wall at each JXF_BC_WALL face             */
  double  dirichlet[6][5];          /* (rho, u, v, w, p) prescribed at each JXF_BC_DIRICHLET face */
  /* ref: active_physics/is_volume_force + forcings/gravity (source_term_solver.py:163-186, space_solver.py:378-384):
   * rhs(rho u_i) += g_i rho, rhs(E) += g . (rho u), from the stage's conservatives */
  int32_t volume_force;             /* 0/1                                                        */
  int32_t no_convective_flux;       /* 1: active_physics/is_convective_flux = false (dissipative fluxes only,
                                       space_solver.py:517-545); the stage then runs unfused              */
  double  gravity[3];
  /* ref: conservatives/positivity/flux_limiter + flux_partition (solvers/positivity/limiter_flux.py:146-330,
   * space_solver.py:532-543): faces whose flux would drive a neighbour cell's density / pressure below eps under the
   * pseudo-integration with the physical time step fall back to the first-order HLLC flux. The sweeps read the time
   * step from the device scalar of jxf_stage / jxf_step_fused; jxf_sweep / jxf_compute_rhs use jxf_bind_timestep. */
  int32_t flux_limiter;             /* JXF_FLUXLIM_*                                              */
  int32_t flux_partition;           /* JXF_PARTITION_*                                            */
  /* ref: conservatives/convective_fluxes/convective_solver (solvers/convective_fluxes/__init__.py:7-12).  With
   * JXF_SOLVER_FLUX_SPLITTING (flux_splitting_scheme.py:62-111) `stencil` is flux_splitting/reconstruction_stencil,
   * `flux_splitting` the eigenvalue choice (eigendecomposition.py:664-689), the frozen state is ARITHMETIC; recon,
   * riemann, signal_speed and the positivity limiters are not read. */
  int32_t convective_solver;        /* JXF_SOLVER_*                                               */
  int32_t flux_splitting;           /* JXF_FS_*                                                   */
  int32_t frozen_state;             /* JXF_FROZEN_*: state the characteristic decompositions are frozen at */
} jxf_config;

enum { JXF_FLUXLIM_NONE = 0, JXF_FLUXLIM_SIMPLE = 1, JXF_FLUXLIM_NASA = 2 };
enum { JXF_PARTITION_UNIFORM = 0, JXF_PARTITION_CELLSIZE = 1 };

typedef struct jxf_solver* jxf_handle;

const char* jxf_last_error(void);
int jxf_version(void);

/* Builds the immutable launch plan for one block. Host-only; touches no device memory. */
int jxf_create(const jxf_config* cfg, jxf_handle* out);
int jxf_destroy(jxf_handle h);

/* Number of fp64 elements of a halo'd (5,X,Y,Z) buffer / of the interior-only rhs buffer. */
int64_t jxf_field_elems(jxf_handle h);
int64_t jxf_rhs_elems(jxf_handle h);
/* Number of RK stages of the configured integrator. */
int jxf_num_stages(jxf_handle h);

/* ref: SpaceSolver.compute_rhs (solvers/space_solver.py:151-453), convective single-phase branch:
 * rhs = 0.0 + rhs_x + rhs_y + rhs_z over the active axes.  prims: in, rhs: out. */
int jxf_compute_rhs(jxf_handle h, const double* prims, double* rhs, void* stream);

/* The device scalar holding the physical time step size that jxf_sweep / jxf_sweep_range / jxf_compute_rhs hand to
 * the positivity flux limiter (ref: the physical_timestep_size argument of SpaceSolver.compute_rhs,
 * space_solver.py:151-164). Only needed when jxf_config.flux_limiter is set; the pointer is stored, not read. */
int jxf_bind_timestep(jxf_handle h, const double* dt);

/* ref: SpaceSolver.compute_rhs_xi (space_solver.py:456-674): one axis.
 * accumulate=0: rhs = 0.0 + rhs_axis ; accumulate=1: rhs += rhs_axis. */
int jxf_sweep(jxf_handle h, int axis, const double* prims, double* rhs, int accumulate, void* stream);

/* One fused RK stage = compute_rhs + TimeIntegrator.perform_stage_integration
 * (time_integrator.py:108-227, RK3.py:27-62) + get_primitives_from_conservatives
 * (equation_manager.py:164-171) + outer face halo fill (halo_manager.py:146-234).
 *   prims_in   : primitives at stage entry (with halos)                  [in]
 *   prims_out  : primitives after the stage, must not alias prims_in     [out]
 *   cons_in    : conservatives at stage entry                            [in]
 *   cons_n     : conservatives at step entry U^n (unused for stage 0)    [in]
 *   cons_out   : conservatives after the stage; may alias cons_in/cons_n [out]
 *   rhs_scratch: interior-only scratch for the partial sums of the first
 *                sweeps (may be NULL with one active axis)               [scratch]
 *   dt_dev     : device pointer to the step's dt                         [in]
 *   red_dev    : device pointer to 3 doubles {max sum(|u_i|+c), min rho, min p};
 *                updated (max/min-combined) when `reduce` != 0; caller resets via jxf_reduce_reset.
 *   fill_halo  : 1 = the outer-BC face halos (PERIODIC/SYMMETRY/ZEROGRADIENT faces) of prims_out and
 *                cons_out are written by the stage itself (fused into the last sweep's epilogue);
 *                0 = halos are left untouched.  Faces marked JXF_BC_NEIGHBOR are never touched. */
int jxf_stage(jxf_handle h, int stage, const double* prims_in, double* prims_out,
              const double* cons_in, const double* cons_n, double* cons_out,
              double* rhs_scratch, const double* dt_dev, double* red_dev,
              int reduce, int fill_halo, void* stream);

/* The same stage, starting at the `first_axis_index`-th ACTIVE axis (0 = whole stage).  With
 * jxf_sweep_range this lets the host run the first sweep in pieces -- interior cells while the
 * inter-block halo exchange of the previous stage is still in flight, the cells next to shared
 * faces afterwards (ref: the halo update at simulation_manager.py:963 precedes the next
 * compute_rhs at :796; only the sweep ALONG an axis reads that axis' halos). */
int jxf_stage_tail(jxf_handle h, int stage, int first_axis_index, const double* prims_in, double* prims_out,
                   const double* cons_in, const double* cons_n, double* cons_out,
                   double* rhs_scratch, const double* dt_dev, double* red_dev,
                   int reduce, int fill_halo, void* stream);

/* jxf_sweep restricted to the cells [lo, hi) along `axis` (strided axes only; the contiguous
 * axis accepts only the full range). */
int jxf_sweep_range(jxf_handle h, int axis, int lo, int hi, const double* prims, double* rhs,
                    int accumulate, void* stream);

/* One whole time step on a single block (ref: SimulationManager._do_integration_step,
 * simulation_manager.py:536-668, and do_runge_kutta_stages :670-1077): all RK stages with
 * local halo fill, the reductions on the last stage, then jxf_finish_step.
 * State at entry: (prims_a, cons_a).  cons ends in cons_a; prims end in prims_a if the
 * return value is 0, in prims_b if it is 1.  Negative return = error.  No host sync. */
int jxf_step_fused(jxf_handle h, double* prims_a, double* prims_b, double* cons_a, double* cons_b,
                   double* rhs_scratch, double* dt_dev, double* time_dev, double* red_dev,
                   double* info_dev, int fill_halo, void* stream);

/* ref: HaloManager.perform_halo_update_material (halos/halo_manager.py:146-234) for
 * PERIODIC / SYMMETRY / ZEROGRADIENT faces (halos/outer/material.py:868-894); cons halos are
 * recomputed from prim halos (:248-250).  Faces marked JXF_BC_NEIGHBOR are skipped. In place.
 * With the viscous or heat flux active the EDGE halos are filled too (halo_manager.py:119-129,
 * :193-199 -> jxf_halo_fill_edges). */
int jxf_halo_fill(jxf_handle h, double* prims, double* cons, void* stream);

/* ref: BoundaryConditionMaterial.edge_halo_update / compute_edge_halos (halos/outer/material.py:289-383)
 * with the type combination of boundary_condition.py:128-179 and the retrieve table :607-655: per edge
 * (pair of faces) the FIRST face that is PERIODIC or SYMMETRY decides -- PERIODIC: copy across the
 * domain, SYMMETRY: mirror and negate that face's normal velocity -- otherwise the mean of the two
 * adjacent halo regions; cons recomputed.  Needs the face halos filled.  In place. */
int jxf_halo_fill_edges(jxf_handle h, double* prims, double* cons, void* stream);

/* ref: SourceTermSolver.compute_viscous_flux_xi + compute_heat_flux_xi (solvers/source_term_solver.py:188-345,
 * :405-470, :503-582; CENTRAL4 stencils/derivative/deriv_face_4.py, deriv_center_4.py,
 * stencils/reconstruction/central/central_4.py) folded into the flux divergence as
 * space_solver.py:567-599 does: rhs (+)= (1/dx) (Fd_{i-1/2} - Fd_{i+1/2}), Fd = (0, -tau, -u.tau + q).
 * accumulate=0 writes (mass row = 0), accumulate=1 adds.  Needs face AND edge halos of prims. */
int jxf_dissipative_sweep(jxf_handle h, int axis, const double* prims, double* rhs, int accumulate, void* stream);

/* ref: MaterialManager.get_temperature -> IdealGas.get_temperature (ideal_gas.py:64-65): T = p/(rho R) on
 * the whole halo'd buffer; temperature: (X, Y, Z) doubles. */
int jxf_temperature(jxf_handle h, const double* prims, double* temperature, void* stream);

/* ref: EquationManager.get_primitives_from_conservatives / get_conservatives_from_primitives
 * (equation_manager.py:164-171, 93-101) over the whole halo'd buffer. */
int jxf_prims_from_cons(jxf_handle h, const double* cons, double* prims, void* stream);
int jxf_cons_from_prims(jxf_handle h, const double* prims, double* cons, void* stream);

/* ref: compute_time_step_size (time_integration/time_step_size.py:15-157) and the min rho / min p
 * logging reductions (solvers/positivity/positivity_handler.py:245-254).
 * jxf_reduce      : red_dev <- combine(red_dev, reductions over the interior of prims)
 * jxf_reduce_reset: red_dev <- {0, +inf, +inf}
 * jxf_finish_step : dt_dev <- CFL*min(dx_min/(red[0]+eps), 3/14 dx^2/(max nu+eps), 0.1 dx^2/(max alpha+eps))
 *                   (diffusive limits only with the viscous / heat flux, time_step_size.py:111-135; or
 *                   fixed_dt); time_dev += dt_used;
 *                   info_dev[0..2] <- red; red_dev reset.  All on device, no host sync. */
int jxf_reduce(jxf_handle h, const double* prims, double* red_dev, void* stream);
int jxf_reduce_reset(jxf_handle h, double* red_dev, void* stream);
int jxf_finish_step(jxf_handle h, double* red_dev, double* dt_dev, double* time_dev,
                    double* info_dev, void* stream);

/* ref: TimeIntegrator.perform_stage_integration (time_integration/time_integrator.py:108-227) with
 * prepare_buffer_for_integration / integrate of RungeKutta3 (RK3.py:49-60), RK2.py, euler.py:
 * whole buffer  U <- a_s U + b_s U^n  (stage > 0), then interior  U += (dt m_s) rhs.
 * Stand-alone form of what jxf_stage fuses into its last sweep; dt is a host scalar like the
 * reference's physical_timestep_size.  cons_out may alias cons. */
int jxf_integrate_stage(jxf_handle h, int stage, const double* cons, const double* cons_n, const double* rhs,
                        double dt, double* cons_out, void* stream);

/* Inter-block face exchange helpers (ref: halos/inner/material.py:30-93).
 * pack  : copies the `nh` interior layers adjacent to `face` of prims into a dense slab
 *         (5, nh, T1, T2) (transverse extents = interior), ready for ncclSend.
 * unpack: writes a received slab into the halo layers of `face` of prims and recomputes the
 *         conservatives there (:83-88). */
int64_t jxf_face_slab_elems(jxf_handle h, int face);
int jxf_pack_face(jxf_handle h, int face, const double* prims, double* slab, void* stream);
int jxf_unpack_face(jxf_handle h, int face, const double* slab, double* prims, double* cons, void* stream);

/* The same with the transverse range widened over the nh halo cells of a transverse axis (ref: the inter-block
 * EDGE halo update, halos/inner/halo_communication.py edge tables, needed by the dissipative fluxes).
 * ext_mask: bit 0 / 1 = low / high side of the slower transverse axis, bit 2 / 3 = of the faster one.
 * Exchanging the faces axis by axis -- x faces widened over the PHYSICAL y/z face halos, y faces over all x
 * halos and the physical z halos, z faces over all x and y halos -- fills every edge halo next to a shared
 * face with what a single-block halo update would put there. */
int64_t jxf_face_slab_elems_ext(jxf_handle h, int face, int ext_mask);
int jxf_pack_face_ext(jxf_handle h, int face, int ext_mask, const double* prims, double* slab, void* stream);
int jxf_unpack_face_ext(jxf_handle h, int face, int ext_mask, const double* slab, double* prims, double* cons,
                        void* stream);

/* The same for the `layers` (1 .. nh) cell layers next to the face only.  The convective stencils read 3 cells
 * beyond a face (ref: spatial_reconstruction.py:21-41 with the WENO5 6-cell window), the reference ships all nh
 * (halos/inner/material.py:74-88): exchanging 3 between RK stages moves 40 % less data; the layers farther out keep
 * their previous values until an nh-layer exchange completes them (BlockRuntime.complete_halos). */
int64_t jxf_face_slab_elems_n(jxf_handle h, int face, int ext_mask, int layers);
int jxf_pack_face_n(jxf_handle h, int face, int ext_mask, int layers, const double* prims, double* slab, void* stream);
int jxf_unpack_face_n(jxf_handle h, int face, int ext_mask, int layers, const double* slab, double* prims, double* cons,
                      void* stream);

/* Boundary DATA of one outer face, applied on top of the face's base rule (bc[face]) by jxf_halo_fill and by the fused
 * halo images of jxf_stage: what the reference's NEUMANN, SIMPLE_INFLOW, SIMPLE_OUTFLOW boundaries, DIRICHLET boundaries
 * with space-dependent primitives_callable, WALL boundaries with a space-dependent wall_velocity_callable and faces with
 * several types prescribe (ref: halos/outer/material.py:473-520, :732-798, :825-866, :966-1050, :121-277).
 * data_dev: (5, n1, n2) doubles over the face's transverse INTERIOR cells (the two other axes in increasing order, the
 * later one fastest), the same for all nh halo layers; ops: 2 bits per variable v at bits 2v: 0 keep the base rule's
 * value, 1 replace it by data_v, 2 add data_v to it; mask_dev: (n1, n2) bytes or NULL -- apply only where != 0.
 * The caller owns both arrays and keeps them alive; ops = 0 clears the face. */
int jxf_set_face_data(jxf_handle h, int face, int ops, const double* data_dev, const unsigned char* mask_dev);

/* Peer-memory halo exchange (multi-GPU; replaces the ppermute of the face slabs, ref: halos/inner/material.py:30-93,
 * halos/inner/halo_communication.py:35-77).  The caller maps the neighbours' field buffers into this process (CUDA IPC)
 * and, before a stage, names per NEIGHBOR face the neighbour's OUTPUT buffers of that stage (jxf_set_peer_halo; NULL,
 * NULL clears): jxf_stage's fused epilogue then stores the images of the cells within nh of the face -- the neighbour's
 * halo cells, primitives and recomputed conservatives, all nh layers -- straight into the neighbour's memory.
 * jxf_peer_signal (after the stage, same stream) publishes `epoch` in each neighbour's flag word
 * (neighbour_flag_slots[face]: peer-mapped pointer to the word the neighbour polls for its opposite face, or NULL);
 * jxf_peer_wait (before the first kernel that reads those halos) spins on this block's own flag words `flags[6]` until
 * every face in `face_mask` shows >= epoch.  Enqueue-only; every rank must signal an epoch before it waits for it. */
/* Mapping a neighbour's buffer (one process per GPU): jxf_peer_export names the device allocation that holds `ptr` --
 * a 64-byte CUDA IPC handle and ptr's byte offset inside that allocation -- for the caller to send to the other process
 * (any transport); jxf_peer_import opens such a handle in the context of the CURRENT device (peer access to the owner's
 * device is enabled by the driver, NVLink / NVSwitch) and returns the mapped address of the same byte.  An allocation is
 * opened once per process and device (cached); jxf_peer_release unmaps everything this process imported. */
int jxf_peer_export(const void* ptr, unsigned char* handle_out /* [64] */, int64_t* offset_out);
int jxf_peer_import(const unsigned char* handle /* [64] */, int64_t offset, void** ptr_out);
int jxf_peer_release(void);
int jxf_set_peer_halo(jxf_handle h, int face, double* peer_prims_out, double* peer_cons_out);
int jxf_peer_signal(jxf_handle h, int64_t* const* neighbour_flag_slots, int64_t epoch, void* stream);
int jxf_peer_wait(jxf_handle h, const int64_t* flags, int face_mask, int64_t epoch, void* stream);

/* One RK stage on THREE full-size buffers: `prims` is updated IN PLACE (no ping-pong buffer) and the rhs accumulator is
 * two slabs of `slab_planes` x planes (jxf_rhs_slab_elems doubles each, stored back to back in `rhs_slabs`).  The block
 * is processed slab by slab with the x sweep running one slab ahead (see the definition).  3-D blocks, convective flux
 * only.  Same results as jxf_stage (same kernels and arithmetic; the launch plan differs).  cons_out may alias cons_in.
 * Footprint at 1024^3: 3 x 44.2 GB + 2 x 2.7 GB against 221 GB for jxf_stage's plan.
 * ref: time_integration/RK3.py:27-62, solvers/space_solver.py:266-314. */
int64_t jxf_rhs_slab_elems(jxf_handle h, int slab_planes);
int jxf_stage_inplace(jxf_handle h, int stage, double* prims, const double* cons_in, const double* cons_n,
                      double* cons_out, double* rhs_slabs, int slab_planes, const double* dt_dev, double* red_dev,
                      int reduce, int fill_halo, void* stream);

/* Launch accounting and optional per-kernel timing (bench / roofline evidence).
 * Kinds: 0..2 = sweep along axis 0..2 writing rhs; 3..5 = sweep along axis 0..2 with the fused
 * RK-stage epilogue; 6 = halo fill; 7 = other (transforms, reductions, pack/unpack); 8 = dissipative
 * (viscous + heat flux) sweeps.
 * With profiling enabled every launch is bracketed by cudaEventRecord on the launch stream
 * (up to 4096 launches between reads).  jxf_profile_read synchronises on the recorded events and
 * returns, per kind, the summed device time in ms, the number of timed launches and the number of
 * launches issued since the last reset. */
#define JXF_PROFILE_KINDS 9
#define JXF_PROFILE_HALO 6
#define JXF_PROFILE_OTHER 7
#define JXF_PROFILE_DISSIPATIVE 8
int jxf_profile_enable(jxf_handle h, int enable);
int jxf_profile_read(jxf_handle h, double* ms_sum, int64_t* timed, int64_t* launches, int reset);

/* Test hook (no device work; callable without a GPU): the kernel instantiation and the face-flux option word the
 * handle's configuration selects -- RECON / RIEMANN template parameters of the sweep kernels and the packed options of
 * numerics.cuh face_flux -- so that the host simulation of the device functions can be checked to use the same ones. */
int jxf_debug_dispatch(jxf_handle h, int axis, int* recon_template, int* riemann_template, int* option_word);

/* Test hook: the per-face device function (reconstruction + Riemann flux,
 * ref: HighOrderGodunov.compute_flux_xi, high_order_godunov.py:117-231) on caller-supplied
 * 6-cell windows.  windows: (n, 5, 6) doubles, flux: (n, 5) doubles, both on the device.
 * `recon` = JXF_RECON_* + 2 * JXF_STENCIL_*. */
int jxf_debug_face_flux(int axis, int recon, int riemann, const double* windows, int64_t n,
                        double gamma, double* flux, void* stream);

/* Test hook: accuracy of the reciprocal / rsqrt building blocks.  x: n positive doubles;
 * out: (n, 4) = {MUFU rcp seed, MUFU rsqrt seed, rcp_fast(x), rsqrt_fast(x)}. */
int jxf_debug_math(const double* x, int64_t n, double* out, void* stream);

/* Device FP64 FMA throughput probe for the roofline denominator (bench only):
 * runs `iters` dependent-chain DFMAs x 8 chains per thread on a full grid; returns the
 * number of DFMA issued in *n_fma; time it with events around the call. */
int jxf_fp64_probe(double* scratch, int iters, int64_t* n_fma, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* JXF_B200_H */
