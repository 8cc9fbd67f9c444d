"""BlockRuntime: the device state of this rank's block and the step/stage drivers
that the API-level managers (InitializationManager, SimulationManager and its
SpaceSolver / TimeIntegrator / HaloManager views) share.

Memory plan per block (fp64): two primitive buffers (stage ping-pong: a sweep reads
its neighbours' primitives, so a stage cannot update them in place), two
conservative buffers (U and U^n swap roles, no per-step copy), one interior-only
rhs accumulator, and one send + one recv slab per shared face.
"""
from __future__ import annotations

import os
import weakref
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .domain_information import FACES
from .engine import BlockConfig, BlockSolver
from .parallel import FACE_ID, ParallelContext

EPS = float(np.finfo(np.float64).eps)


class BlockRuntime:
    _cache: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()

    @classmethod
    def get(cls, input_manager, parallel: Optional[ParallelContext] = None) -> "BlockRuntime":
        rt = cls._cache.get(input_manager)
        if rt is None:
            if parallel is None:
                parallel = ParallelContext.from_environment(input_manager.domain_information)
            rt = cls(input_manager, parallel)
            cls._cache[input_manager] = rt
        elif parallel is not None and parallel is not rt.parallel and (
                parallel.rank, parallel.world_size) != (rt.parallel.rank, rt.parallel.world_size):
            raise ValueError("BlockRuntime.get: a runtime for this InputManager already exists with a different "
                             "ParallelContext (rank / world size)")
        return rt

    def __init__(self, input_manager, parallel: ParallelContext):
        di = input_manager.domain_information
        num = input_manager.numerical_setup
        case = input_manager.case_setup
        cf = num.conservatives.convective_fluxes
        god, fs = cf.godunov, cf.flux_splitting
        ti = num.conservatives.time_integration
        self.parallel = parallel
        self.domain_information = di
        self.bc_global = dict(case.boundary_condition_setup)
        self.bc_block = parallel.block_boundary_types(self.bc_global)
        self.neighbors = parallel.neighbors(self.bc_global)
        self._cell_sizes = tuple(di.cell_sizes)           # as the reference forms them (domain_information.py:290)
        self.cfg = BlockConfig(
            cells=tuple(di.device_number_of_cells),
            inv_dx=tuple(float(x) for x in di.one_cell_sizes),
            dx_min=di.smallest_cell_size,
            gamma=case.material_setup.specific_heat_ratio,
            bc=self._kernel_boundary_types(),
            nh=di.nh_conservatives,
            convective_solver=cf.convective_solver,
            flux_splitting=fs.flux_splitting if fs is not None else "ROE",
            frozen_state=(god if god is not None else fs).frozen_state,
            recon=god.reconstruction_variable if god is not None else "CHAR-PRIMITIVE",
            stencil=god.reconstruction_stencil if god is not None else fs.reconstruction_stencil,
            riemann=god.riemann_solver if god is not None else "HLLC",
            signal_speed=god.signal_speed if god is not None else "EINFELDT",
            integrator=ti.integrator,
            cfl=ti.CFL,
            fixed_dt=float(ti.fixed_timestep) if ti.fixed_timestep else 0.0,
            is_viscous_flux=num.active_physics.is_viscous_flux,
            is_heat_flux=num.active_physics.is_heat_flux,
            is_viscous_heat_production=num.active_physics.is_viscous_heat_production,
            dynamic_viscosity=case.material_setup.transport.dynamic_viscosity,
            bulk_viscosity=case.material_setup.transport.bulk_viscosity,
            thermal_conductivity_model=case.material_setup.transport.thermal_conductivity_model,
            thermal_conductivity=case.material_setup.transport.thermal_conductivity,
            prandtl_number=case.material_setup.transport.prandtl_number,
            gas_constant=case.material_setup.specific_gas_constant,
            is_interpolation_limiter=num.conservatives.positivity.is_interpolation_limiter,
            limit_velocity=num.conservatives.positivity.limit_velocity,
            flux_limiter=num.conservatives.positivity.flux_limiter,
            flux_partition=num.conservatives.positivity.flux_partition,
            dirichlet=self._dirichlet_constants(case, di, parallel),
            wall_velocity=self._wall_constants(case, di, parallel),
            is_volume_force=num.active_physics.is_volume_force,
            is_convective_flux=num.active_physics.is_convective_flux,
            gravity=tuple(case.gravity),
        )

        self.solver = BlockSolver(self.cfg)
        s = self.solver
        self.device = s.device
        # boundary DATA of this block's outer faces (space-dependent DIRICHLET / WALL values, NEUMANN increments,
        # SIMPLE_INFLOW / SIMPLE_OUTFLOW values, the DIRICHLET parts of multi-type faces): device arrays over the face's
        # transverse cells that the halo kernel and the fused halo images apply on top of the face's base rule
        # (jxf_set_face_data); the tensors are owned here
        self.face_data = self._make_face_data()
        if self.cfg.is_dissipative and any(isinstance(f, tuple) for f in self._host_faces):
            raise NotImplementedError("several boundary types on one face together with the viscous / heat flux (edge "
                                      "halos next to such a face) are not implemented on the B200 path")
        for f, (ops, data, mask) in self.face_data.items():
            s.set_face_data(FACE_ID[f], ops, data, mask)
        self._host_halo = False                # (kept for the memory-plan / graph conditions: nothing is host-applied)
        self.stages = s.stages
        # Memory plan.  "pingpong" (default): 2 primitive + 2 conservative buffers + a full-size rhs accumulator.
        # "inplace": ONE primitive buffer updated in place + 2 conservative buffers + two slab-sized rhs accumulators
        # (jxf_stage_inplace) -- 3 full-size buffers instead of 5, for blocks the default plan cannot hold (1024^3 on one
        # GPU).  JXF_MEMORY_PLAN=pingpong|inplace|auto; auto picks in place when the default plan exceeds the free memory.
        self.memory_plan = self._pick_memory_plan()
        self.slab_planes = int(os.environ.get("JXF_SLAB_PLANES", "64"))
        if self.memory_plan == "inplace":
            p = s.new_field(EPS)
            self.prims = [p, p]                                 # one buffer: the stage updates it in place
            self.rhs = None
            self.rhs_slabs = s.new_rhs_slabs(self.slab_planes)
        else:
            self.prims = [s.new_field(EPS), s.new_field(EPS)]   # helper_functions.py:21-60: eps fill
            self.rhs = s.new_rhs() if (len(s.active) > 1 or self.cfg.is_dissipative) else None
        self.cons = [s.new_field(EPS), s.new_field(EPS)]        # cons[0] = U / U^n, cons[1] = stage scratch
        self.cur = 0
        self.red = s.new_red()
        self.info = s.new_scalars(3)
        self.time = s.new_scalars(1, 0.0)
        self.dt = s.new_scalars(1, 0.0)
        if self.cfg.flux_limiter:           # sweeps issued outside jxf_stage (overlap pieces, compute_rhs) read this dt
            s.bind_timestep(self.dt)
        self._sign = torch.tensor([1.0, -1.0, -1.0], dtype=torch.float64, device=self.device)
        # With the dissipative fluxes the exchange also carries EDGE halos: faces go axis by axis, each slab widened
        # over the transverse halos that are already complete (ext_mask, see jxf_pack_face_ext)
        self.ext_mask = {f: (self._edge_ext_mask(f) if self.cfg.is_dissipative else 0) for f in self.neighbors}
        self.send = {f: torch.empty(s.face_slab_elems(FACE_ID[f], self.ext_mask[f]), dtype=torch.float64,
                                    device=self.device) for f in self.neighbors}
        self.recv = {f: torch.empty_like(self.send[f]) for f in self.neighbors}
        # Layers shipped per shared face between RK stages.  The convective stencils read REACH = 3 cells beyond a face,
        # the reference ships all nh (halos/inner/material.py:74-88).  stage_layers < nh leaves the outer halo layers
        # stale until complete_halos() -- which every API call that hands buffers to the user runs -- so what the user
        # sees is bit-identical to a full exchange.  The dissipative stencils reach 2 + 2 cells: full exchange there.
        self.stage_layers = self.cfg.nh
        if not self.cfg.is_dissipative and os.environ.get("JXF_EXCHANGE_LAYERS", "") != "full":
            self.stage_layers = min(self.cfg.nh, int(os.environ.get("JXF_EXCHANGE_LAYERS", self.REACH)))
        self._halos_partial = False
        self._comm_prof = None                    # comm_profile(True): lists of (name, start event, stop event)
        # inter-block exchange runs on its own stream and overlaps the first sweep of the next stage
        self.overlap = (bool(self.neighbors) and os.environ.get("JXF_OVERLAP", "1") != "0"
                        and not self.cfg.is_dissipative)
        # high priority: the pack / unpack kernels of the exchange share the SMs with the overlapped sweep (which
        # fills every SM); with a high-priority stream their CTAs take the slots that free up first
        prio = -1 if os.environ.get("JXF_COMM_PRIORITY", "1") != "0" else 0
        self.comm_stream = torch.cuda.Stream(device=self.device, priority=prio) if self.neighbors else None
        self._pending = None                      # event: halos of the current primitives are complete
        # Peer-memory halo exchange (JXF_PEER_HALO=1): the neighbours' field buffers are mapped into this process (CUDA
        # IPC over NVLink) and the fused epilogue stores the halo images of shared faces straight into them; the
        # exchange between stages shrinks to one flag word per shared face (jxf_peer_signal / jxf_peer_wait).
        self.peer = None
        if (self.neighbors and os.environ.get("JXF_PEER_HALO", "0") == "1" and not self.cfg.is_dissipative
                and self.memory_plan == "pingpong"):
            self._setup_peer()
        first = s.active[0]
        self._first_axis = first
        self._first_strided = len(s.active) > 1   # the contiguous (last active) axis takes no partial ranges
        self._first_split = any(f in self.neighbors for f in (FACES[2 * first], FACES[2 * first + 1]))

    def _pick_memory_plan(self) -> str:
        want = os.environ.get("JXF_MEMORY_PLAN", "auto").lower()
        can = (len(self.solver.active) == 3 and not self.cfg.is_dissipative and self.cfg.is_convective_flux
               and not self._host_halo)
        if want == "inplace":
            if not can:
                raise NotImplementedError("JXF_MEMORY_PLAN=inplace: 3-D convective-only blocks without host-applied boundaries")
            return "inplace"
        if want == "pingpong" or not can:
            return "pingpong"
        field = int(np.prod(self.cfg.shape)) * 8
        need = 4 * field + int(np.prod(self.cfg.rhs_shape)) * 8
        try:
            free, _ = torch.cuda.mem_get_info(self.device)
        except Exception:
            return "pingpong"
        return "inplace" if need > 0.92 * free else "pingpong"

    # -- boundaries with data beyond the kernels' constants -----------------------
    # DIRICHLET with space-dependent data, NEUMANN, SIMPLE_INFLOW, SIMPLE_OUTFLOW, space-dependent WALL velocities,
    # multi-type faces: the kernels are configured with the base rule of the face (constants / ZEROGRADIENT / WALL at
    # rest) and apply the prescribed data -- device arrays over the face's transverse cells -- on top of it, in the halo
    # kernel and in the fused halo images alike (jxf_set_face_data).  Multi-block runs included.
    KERNEL_TYPE = {"NEUMANN": "ZEROGRADIENT", "SIMPLE_INFLOW": "ZEROGRADIENT", "SIMPLE_OUTFLOW": "ZEROGRADIENT"}

    def _kernel_boundary_types(self) -> Dict[str, str]:
        """The boundary types the kernels are configured with: NEUMANN / SIMPLE_* faces are ZEROGRADIENT there (the source
        cell of all three is the last interior cell, boundary_condition.py:580-595)."""
        return {f: self.KERNEL_TYPE.get(t, t) for f, t in self.bc_block.items()}

    def _dirichlet_constants(self, case, di, parallel) -> Dict[str, Tuple[float, ...]]:
        """Evaluate every primitives_callable of this block's outer faces (halos/outer/material.py:770-790, :825-866,
        :966-1050).  DIRICHLET faces whose values are all constants go to the kernels (jxf_config.dirichlet); everything
        else is kept in self._host_faces = {face: (type, values)}, DIRICHLET faces among them with a finite placeholder
        in the kernels' table."""
        from .input_manager import evaluate_dirichlet_face
        consts, self._host_faces = {}, {}
        for f, values in dict(case.dirichlet_setup).items():
            t = self.bc_block.get(f)
            if t not in ("DIRICHLET", "NEUMANN", "SIMPLE_INFLOW", "SIMPLE_OUTFLOW"):   # a face shared with a neighbour
                continue
            vals = evaluate_dirichlet_face(values, f, di, parallel.rank)
            if t != "DIRICHLET":
                self._host_faces[f] = (t, vals, None)
            elif all(isinstance(v, float) for v in vals):
                consts[f] = vals
            else:
                self._host_faces[f] = (t, vals, None)
                consts[f] = tuple(v if isinstance(v, float) else float(np.ravel(v)[0]) for v in vals)
        # faces with several types: the kernels fill the ZEROGRADIENT / SYMMETRY entry on the whole face, the DIRICHLET
        # entries are written where their bounding_domain holds.  The masks must partition the face (then the order
        # in which the reference applies the entries, material.py:121-277, does not matter).
        from .input_manager import evaluate_bounding_domain
        for f, m in dict(getattr(case, "multi_type_setup", {}) or {}).items():
            if self.bc_block.get(f) not in ("ZEROGRADIENT", "SYMMETRY"):
                continue
            covered = evaluate_bounding_domain(m["kernel_bounding_domain"], f, di, parallel.rank).astype(int)
            for i, (values, dom) in enumerate(m["dirichlet"]):
                mask = evaluate_bounding_domain(dom, f, di, parallel.rank)
                covered = covered + mask.astype(int)
                self._host_faces[(f, i)] = ("DIRICHLET", evaluate_dirichlet_face(values, f, di, parallel.rank), mask)
            if not np.all(covered == 1):
                raise NotImplementedError(f"boundary_conditions/{f}: the bounding domains of the face's types must "
                                          "partition the face on the B200 path")
        return consts

    def _wall_constants(self, case, di, parallel) -> Dict[str, Tuple[float, float, float]]:
        """wall_velocity_callable of the WALL faces (halos/outer/material.py:473-520).  Constant velocities go to the
        kernels; a face with a space-dependent component gets velocity 0 there (halo = -u_mirror) and the host adds
        2 u_wall on top (_apply_host_boundaries).  Called after _dirichlet_constants (which creates self._host_faces)."""
        from .input_manager import evaluate_dirichlet_face
        consts = {}
        for f, uvw in dict(case.wall_velocity_setup).items():
            if self.bc_block.get(f) != "WALL":
                continue
            vals = evaluate_dirichlet_face((None,) + tuple(uvw) + (None,), f, di, parallel.rank, "wall_velocity_callable")
            if all(isinstance(v, float) for v in vals[1:4]):
                consts[f] = tuple(vals[1:4])
            else:
                self._host_faces[f] = ("WALL", vals, None)
                consts[f] = (0.0, 0.0, 0.0)
        return consts

    # ops of jxf_set_face_data per variable: 1 = replace by the data, 2 = add the data to the base rule's value
    def _make_face_data(self):
        """{face: (ops, data (5, n1, n2) device tensor, mask (n1, n2) uint8 device tensor or None)} from self._host_faces.
        NEUMANN: the increment (value * upwind sign) * dx of halos/outer/material.py:857-862 added to the ZEROGRADIENT
        copy; SIMPLE_INFLOW: rho, u, v, w replaced, p kept (:966-1022); SIMPLE_OUTFLOW: p replaced (:1024-1050); DIRICHLET:
        all five replaced (:770-790); WALL: 2 u_wall added to -u_mirror (:494-496, the kernels run with u_wall = 0
        there); entries of a multi-type face: replaced inside the union of their bounding domains (:121-277)."""
        out = {}
        for key, (kind, vals, mask) in self._host_faces.items():
            face = key[0] if isinstance(key, tuple) else key
            ax = FACE_ID[face] >> 1
            hi = (FACE_ID[face] & 1) == 0                      # east / north / top
            shape3 = [n if n > 1 else 1 for n in self.cfg.cells]
            shape3[ax] = 1
            tshape = tuple(n for i, n in enumerate(self.cfg.cells) if i != ax)       # (n1, n2), physical order
            data = np.zeros((5,) + tshape)
            ops = 0
            for v, val in enumerate(vals):
                if val is None:
                    continue                               # SIMPLE_INFLOW keeps the copied p, SIMPLE_OUTFLOW rho, u, v, w
                a = np.broadcast_to(np.asarray(val, dtype=np.float64), shape3).reshape(tshape)
                op = 1
                if kind == "NEUMANN":
                    a = a * (-1 if hi else 1) * self._cell_sizes[ax]
                    op = 2
                elif kind == "WALL":
                    if v not in (1, 2, 3):
                        continue
                    a = 2 * a                              # u_halo = 2 u_wall - u_mirror
                    op = 2
                data[v] = a
                ops |= op << (2 * v)
            m = None if mask is None else np.broadcast_to(np.asarray(mask, dtype=bool), shape3).reshape(tshape)
            if face in out:                                # a further DIRICHLET entry of the same multi-type face
                ops0, data0, m0 = out[face]
                assert ops0 == ops and m0 is not None and m is not None
                data = np.where(m[None], data, data0)
                m = m | m0
            out[face] = (ops, data, m)
        dev = {}
        for face, (ops, data, m) in out.items():
            dev[face] = (ops, torch.as_tensor(np.ascontiguousarray(data), dtype=torch.float64).to(self.device),
                         None if m is None else torch.as_tensor(np.ascontiguousarray(m.astype(np.uint8))).to(self.device))
        return dev

    # -- views ------------------------------------------------------------
    @property
    def primitives(self) -> torch.Tensor:
        return self.prims[self.cur]

    @property
    def conservatives(self) -> torch.Tensor:
        return self.cons[0]

    def temperature(self, prims: torch.Tensor) -> Optional[torch.Tensor]:
        """material_manager.get_temperature on the halo'd buffer (simulation_manager.py:971-974); None unless the
        viscous / heat flux is active (equation_information.is_compute_temperature)."""
        if not self.cfg.is_dissipative:
            return None
        return self.solver.temperature(prims)

    def adopt(self, primitives: torch.Tensor, conservatives: torch.Tensor):
        """Make externally supplied tensors the current state (copies unless they already are)."""
        self.finish_pending()
        if primitives.data_ptr() != self.primitives.data_ptr():
            self.primitives.copy_(primitives)
        if conservatives.data_ptr() != self.conservatives.data_ptr():
            self.conservatives.copy_(conservatives)

    # -- initialisation -----------------------------------------------------
    def upload_initial_primitives(self, host_interior: np.ndarray) -> Tuple[torch.Tensor, torch.Tensor]:
        """material_fields_initializer.py:148-210 / :590-690: interior <- IC, cons = f(prims) on the
        whole buffer, then the halo update."""
        p = self.prims[self.cur]
        p.fill_(EPS)
        sl = (slice(None),) + self.cfg.interior
        if callable(host_interior):          # a filler of the interior view (InitializationManager: callable user_prime_init)
            host_interior(p[sl])
        elif torch.is_tensor(host_interior):
            p[sl] = host_interior.to(device=self.device, dtype=torch.float64)
        else:
            p[sl] = torch.as_tensor(np.ascontiguousarray(host_interior), dtype=torch.float64).to(self.device)
        self.solver.cons_from_prims(p, self.cons[0])
        self.halo_update(p, self.cons[0])
        return p, self.cons[0]

    def initial_time_step_and_positivity(self, prims: torch.Tensor):
        s = self.solver
        s.reduce_reset(self.red)
        s.reduce(prims, self.red)
        self._allreduce_red()
        s.finish_step(self.red, self.dt, None, self.info)
        info = self.info.cpu().numpy()
        return float(self.dt.item()), float(info[1]), float(info[2])

    # -- halo update ----------------------------------------------------------
    def _edge_ext_mask(self, face: str) -> int:
        """Transverse widening of the slab of a shared face: over ALL halos of the axes exchanged before this one
        (lower axis index), and over the PHYSICAL-boundary halos of the axes exchanged after it."""
        ax = FACE_ID[face] >> 1
        t1 = 1 if ax == 0 else 0
        t2 = 1 if ax == 2 else 2
        mask = 0
        for bit, t in ((0, t1), (2, t2)):
            if self.cfg.cells[t] <= 1:
                continue
            for side, f in ((0, FACES[2 * t + 1]), (1, FACES[2 * t])):      # low side face, high side face
                physical = self.bc_block[f] not in ("NEIGHBOR", "INACTIVE")
                if t < ax or physical:
                    mask |= 1 << (bit + side)
        return mask

    def _halo_update_with_edges(self, prims: torch.Tensor, cons: torch.Tensor, local_done: bool):
        """Dissipative path on several blocks: physical faces + physical edges locally, then the shared faces
        axis by axis with widened slabs (inter-block edge halos; halos/inner/halo_communication.py)."""
        s = self.solver
        if not local_done:
            s.halo_fill(prims, cons)                     # physical faces, then edges between two physical faces
        for ax in range(3):
            faces = {f: nb for f, nb in self.neighbors.items() if FACE_ID[f] >> 1 == ax}
            if not faces:
                continue
            for f in faces:
                s.pack_face(FACE_ID[f], prims, self.send[f], self.ext_mask[f])
            for r in self.parallel.exchange(faces, self.send, self.recv):
                r.wait()
            for f in faces:
                s.unpack_face(FACE_ID[f], self.recv[f], prims, cons, self.ext_mask[f])

    # -- communication timing (bench `comm` object) ------------------------------
    def comm_profile(self, on: bool = True):
        """Record CUDA events around pack / NCCL / unpack (communication stream) and around the compute stream's wait
        for the exchange; read with comm_profile_read()."""
        self._comm_prof = [] if on else None

    def _tick(self, name):
        """Context manager timing a span on the CURRENT stream when profiling is on."""
        import contextlib
        if self._comm_prof is None:
            return contextlib.nullcontext()
        rt = self

        class Span:
            def __enter__(self_):
                self_.a = torch.cuda.Event(enable_timing=True)
                self_.b = torch.cuda.Event(enable_timing=True)
                self_.a.record()

            def __exit__(self_, *exc):
                self_.b.record()
                rt._comm_prof.append((name, self_.a, self_.b))
        return Span()

    def comm_profile_read(self, reset: bool = True):
        """-> {name: total ms} over the spans recorded so far (synchronises)."""
        out = {}
        if self._comm_prof:
            torch.cuda.synchronize()
            for name, a, b in self._comm_prof:
                out[name] = out.get(name, 0.0) + a.elapsed_time(b)
            if reset:
                self._comm_prof = []
        return out

    def halo_update(self, prims: torch.Tensor, cons: torch.Tensor, local_done: bool = False, layers: Optional[int] = None):
        """halo_manager.py:146-234: inter-block faces (inner/material.py:30-93) then outer BCs.  `layers`: cell layers
        shipped per shared face (default: all nh)."""
        s = self.solver
        if self.neighbors and self.cfg.is_dissipative:
            return self._halo_update_with_edges(prims, cons, local_done)
        if self.neighbors:
            nl = int(layers or self.cfg.nh)
            full = nl == self.cfg.nh
            cut = {f: (self.send[f] if full else self.send[f][:s.face_slab_elems(FACE_ID[f], 0, nl)]) for f in self.neighbors}
            got = {f: (self.recv[f] if full else self.recv[f][:cut[f].numel()]) for f in self.neighbors}
            with self._tick("pack"):
                for f in self.neighbors:
                    s.pack_face(FACE_ID[f], prims, cut[f], 0, nl)
            with self._tick("nccl"):
                reqs = self.parallel.exchange(self.neighbors, cut, got)
                for r in reqs:
                    r.wait()
            with self._tick("unpack"):
                for f in self.neighbors:
                    s.unpack_face(FACE_ID[f], got[f], prims, cons, 0, nl)
            self._halos_partial = not full
        if not local_done:
            s.halo_fill(prims, cons)

    def _allreduce_red(self):
        if self.parallel.is_parallel:
            # order the collective after an in-flight P2P halo exchange (they may use different NCCL communicators on
            # different streams; concurrent communicators with rank-dependent launch order can deadlock)
            self.finish_pending()
            buf = self.red * self._sign
            self.parallel.allreduce_max(buf)
            self.red.copy_(buf * self._sign)

    # -- stepping ---------------------------------------------------------------
    def set_time_control(self, time: float, dt: float):
        self.time.fill_(float(time))
        self.dt.fill_(float(dt))

    # stencil reach: the rhs of a cell needs the fluxes of its two faces, whose windows span 3 cells
    # on either side, so only cells within 3 of a shared face read exchanged halos
    REACH = 3

    def _setup_peer(self):
        """Map the neighbours' four field buffers and flag words into this process: every rank exports the CUDA-IPC
        handle of each buffer's allocation (jxf_peer_export), the handles travel through the process group, and each
        rank opens its neighbours' under ITS OWN device (jxf_peer_import), which enables NVLink peer access."""
        import torch.distributed as dist
        from .parallel import OPPOSITE
        s = self.solver
        self.peer_flags = torch.zeros(8, dtype=torch.int64, device=self.device)
        torch.cuda.synchronize(self.device)
        mine = [s.peer_export(t) for t in (self.prims[0], self.prims[1], self.cons[0], self.cons[1], self.peer_flags)]
        gathered = [None] * self.parallel.world_size
        dist.all_gather_object(gathered, mine, group=self.parallel.group)
        with torch.cuda.device(self.device):
            self._peer_ptrs = {}
            for nb in set(self.neighbors.values()):
                assert nb != self.parallel.rank
                self._peer_ptrs[nb] = [s.peer_import(h, off) for h, off in gathered[nb]]
        self.peer = {f: self._peer_ptrs[nb] for f, nb in self.neighbors.items()}
        # the word the neighbour across face f polls for ITS face opposite(f)
        self.peer_slots = [None] * 6
        self.peer_mask = 0
        for f, t in self.peer.items():
            self.peer_slots[FACE_ID[f]] = t[4] + 8 * FACE_ID[OPPOSITE[f]]
            self.peer_mask |= 1 << FACE_ID[f]
        self._epoch = 0
        # prove the mappings before the first stage depends on them: epoch 0 into every neighbour's flag word (a no-op
        # for the protocol), then a synchronise that surfaces an unmapped address HERE rather than inside a sweep
        s.peer_signal(self.peer_slots, 0)
        torch.cuda.synchronize(self.device)
        self.parallel.barrier()

    def _set_peer_outputs(self, last: bool):
        """Before a stage: the neighbours' OUTPUT buffers of this stage (every rank runs the same ping-pong sequence)."""
        for f, t in self.peer.items():
            self.solver.set_peer_halo(FACE_ID[f], t[self.cur ^ 1], t[2] if last else t[3])

    def _start_exchange(self, prims: torch.Tensor, cons: torch.Tensor):
        """Post the inter-block halo exchange of (prims, cons) on the communication stream."""
        if self.peer is not None:      # the stage's epilogue already stored the halos remotely: publish the epoch
            self._epoch += 1
            with self._tick("signal"):
                self.solver.peer_signal(self.peer_slots, self._epoch)
            self._pending = ("peer", self._epoch)
            return
        compute = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(compute)
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ready)
            self.halo_update(prims, cons, local_done=True, layers=self.stage_layers)
            done = torch.cuda.Event()
            done.record(self.comm_stream)
        self._pending = done

    def finish_pending(self):
        """Make the current stream wait for an in-flight halo exchange (before anyone reads halos)."""
        if self._pending is not None:
            with self._tick("wait"):
                if isinstance(self._pending, tuple):       # peer exchange: spin on this block's flag words
                    self.solver.peer_wait(self.peer_flags, self.peer_mask, self._pending[1])
                else:
                    torch.cuda.current_stream().wait_event(self._pending)
            self._pending = None

    def complete_halos(self):
        """Bring every halo layer of the current state up to date (after stages that shipped stage_layers < nh): one
        full exchange.  Run by the API before buffers are handed to the user; the step loop itself never needs it."""
        self.finish_pending()
        if self._halos_partial and self.neighbors:
            self.halo_update(self.prims[self.cur], self.cons[0], local_done=True)
        self._halos_partial = False

    def stage(self, k: int, reduce: bool):
        """One RK stage on the current state (simulation_manager.py:770-1047).  With several blocks the
        exchange of the previous stage's halos may still be in flight: only the sweep ALONG an axis reads
        that axis' halos, so the first sweep runs on the interior cells (or entirely, when its axis is not
        split) before waiting for the exchange."""
        s = self.solver
        last = k == self.stages - 1
        p_in, p_out = self.prims[self.cur], self.prims[self.cur ^ 1]
        c_in = self.cons[0] if k == 0 else self.cons[1]
        c_out = self.cons[0] if last else self.cons[1]
        if self.memory_plan == "inplace":
            self.finish_pending()
            s.stage_inplace(k, p_in, c_in, self.cons[0], c_out, self.rhs_slabs, self.slab_planes, self.dt, self.red,
                            reduce=reduce, fill_halo=True)
            if self.neighbors:          # the in-place stage overwrites what an overlapped sweep would read: no overlap
                self.halo_update(p_out, c_out, local_done=True, layers=self.stage_layers)
            self.cur ^= 1
            return
        args = (p_in, p_out, c_in, self.cons[0], c_out, self.rhs, self.dt, self.red)
        if self.peer is not None:
            self._set_peer_outputs(last)
        if self._pending is not None and self.overlap and self._first_strided:
            ax, n, w = self._first_axis, self.cfg.cells[self._first_axis], self.REACH
            if self._first_split and n > 2 * w:
                s.sweep_range(ax, w, n - w, p_in, self.rhs, accumulate=False)
                self.finish_pending()
                s.sweep_range(ax, 0, w, p_in, self.rhs, accumulate=False)
                s.sweep_range(ax, n - w, n, p_in, self.rhs, accumulate=False)
            else:
                if self._first_split:
                    self.finish_pending()
                s.sweep_range(ax, 0, n, p_in, self.rhs, accumulate=False)
                self.finish_pending()
            s.stage_tail(k, 1, *args, reduce=reduce, fill_halo=True)
        else:
            self.finish_pending()
            s.stage(k, *args, reduce=reduce, fill_halo=True)
        if self.neighbors:
            if self.overlap or self.peer is not None:
                self._start_exchange(p_out, c_out)
            else:
                self.halo_update(p_out, c_out, local_done=True, layers=self.stage_layers)
        self.cur ^= 1

    # -- CUDA graphs ------------------------------------------------------------
    def use_cuda_graph(self, on: bool = True):
        """Replay the single-block step from a captured CUDA graph instead of launching its kernels one by one.
        The step is launch-bound on small grids (Sod-1000: 4 launches, ~27 us of launch latency around ~8 us of
        work); dt, t and the reductions live on the device, so the captured launches are identical every step.
        One graph per primitive-buffer parity (an RK3 step flips the ping-pong), captured lazily."""
        if self.parallel.is_parallel:
            raise NotImplementedError("CUDA-graph replay is implemented for the single-block step")
        self._graphs = {} if on else None

    def _graph_step(self):
        g = self._graphs.get(self.cur)
        if g is None:
            cur0 = self.cur
            # warm up on a side stream (first launches create TMA descriptors / set kernel attributes), then capture
            state = [t.clone() for t in (*self.prims, *self.cons, self.dt, self.time, self.red, self.info)]
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._eager_step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            for t, s0 in zip((*self.prims, *self.cons, self.dt, self.time, self.red, self.info), state):
                t.copy_(s0)
            self.cur = cur0
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._eager_step()
            g = (graph, self.cur)               # cur after one step from cur0
            self.cur = cur0
            for t, s0 in zip((*self.prims, *self.cons, self.dt, self.time, self.red, self.info), state):
                t.copy_(s0)
            self._graphs[cur0] = g
        g[0].replay()
        self.cur = g[1]

    def _eager_step(self):
        a, b = self.prims[self.cur], self.prims[self.cur ^ 1]
        where = self.solver.step_fused(a, b, self.cons[0], self.cons[1], self.rhs, self.dt, self.time, self.red,
                                       self.info)
        self.cur ^= where

    def step(self):
        """One full time step, enqueue-only (no host sync)."""
        if not self.parallel.is_parallel and not self._host_halo and self.memory_plan != "inplace":
            if getattr(self, "_graphs", None) is not None:
                return self._graph_step()
            return self._eager_step()
        for k in range(self.stages):
            self.stage(k, reduce=(k == self.stages - 1))
        self._allreduce_red()
        self.solver.finish_step(self.red, self.dt, self.time, self.info)

    def read_step_scalars(self, complete_halos: bool = False):
        """(time, dt_next, max_speed_sum, min_rho, min_p) -- ONE device->host sync."""
        if complete_halos:
            self.complete_halos()
        self.finish_pending()
        v = torch.cat([self.time, self.dt, self.info]).cpu().numpy()
        return float(v[0]), float(v[1]), float(v[2]), float(v[3]), float(v[4])
