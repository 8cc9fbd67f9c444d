"""Device-side engine of one block: owns the fp64 buffers (torch CUDA tensors used
purely as allocator/stream glue) and drives the C ABI (include/jxf_b200.h).

There is no CPU path here: every method enqueues hand-written sm_100a kernels
from libjxf_b200.so on the current torch CUDA stream.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

FACES = _lib.FACES


@dataclass
class BlockConfig:
    """What jxf_config needs, in the reference's vocabulary."""
    cells: Tuple[int, int, int]                 # interior cells of this block
    inv_dx: Tuple[float, float, float]          # 1/dx per axis (domain_information.py:290)
    dx_min: float
    gamma: float
    bc: Dict[str, str]                          # face -> INACTIVE|PERIODIC|SYMMETRY|ZEROGRADIENT|NEIGHBOR
    nh: int = 5
    recon: str = "CHAR-PRIMITIVE"
    stencil: str = "WENO5-Z"
    riemann: str = "HLLC"
    signal_speed: str = "EINFELDT"
    integrator: str = "RK3"
    cfl: float = 0.5
    fixed_dt: float = 0.0
    # dissipative fluxes (active_physics + material_properties/transport); constants only on this path
    is_viscous_flux: bool = False
    is_heat_flux: bool = False
    is_viscous_heat_production: bool = True
    dynamic_viscosity: float = 0.0
    bulk_viscosity: float = 0.0
    thermal_conductivity_model: str = "CUSTOM"        # CUSTOM | PRANDTL
    thermal_conductivity: float = 0.0
    prandtl_number: float = 1.0
    gas_constant: float = 1.0
    # conservatives/positivity and WALL boundaries
    is_interpolation_limiter: bool = False
    limit_velocity: bool = False
    flux_limiter: Optional[str] = None                 # positivity/flux_limiter: SIMPLE | NASA
    flux_partition: str = "UNIFORM"                    # positivity/flux_partition: UNIFORM | CELLSIZE
    wall_velocity: Dict[str, Tuple[float, float, float]] = field(default_factory=dict)   # face -> constant (u, v, w)
    dirichlet: Dict[str, Tuple[float, ...]] = field(default_factory=dict)               # face -> constant (rho,u,v,w,p)
    is_volume_force: bool = False
    gravity: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    is_convective_flux: bool = True
    convective_solver: str = "GODUNOV"                 # or FLUX-SPLITTING (then `stencil` is the flux_splitting block's)
    flux_splitting: str = "ROE"                        # flux_splitting/flux_splitting: ROE | CLLF | LLF
    frozen_state: str = "ARITHMETIC"                   # godunov/frozen_state or flux_splitting/frozen_state: ARITHMETIC | ROE

    @property
    def is_dissipative(self) -> bool:
        return bool(self.is_viscous_flux or self.is_heat_flux)

    def thermal_conductivity_value(self) -> float:
        """material.py:112-116: CUSTOM value, or PRANDTL cp mu / Pr with cp = gamma/(gamma-1) R (ideal_gas.py:33),
        evaluated in the reference's operation order."""
        if self.thermal_conductivity_model == "CUSTOM":
            return float(self.thermal_conductivity)
        if self.thermal_conductivity_model == "PRANDTL":
            cp = self.gamma / (self.gamma - 1.0) * self.gas_constant
            return cp * float(self.dynamic_viscosity) / float(self.prandtl_number)
        raise NotImplementedError(f"thermal_conductivity model '{self.thermal_conductivity_model}' is not implemented "
                                  "on the B200 path (implemented: CUSTOM, PRANDTL)")

    def to_c(self) -> _lib.JxfConfig:
        def lookup(table, key, what):
            if key not in table:
                raise NotImplementedError(
                    f"{what} '{key}' is a reference option that is not implemented on the B200 path "
                    f"(available: {sorted(table)})")
            return table[key]
        c = _lib.JxfConfig()
        for i in range(3):
            c.n[i] = int(self.cells[i])
            c.inv_dx[i] = float(self.inv_dx[i])
        c.nh = int(self.nh)
        c.dx_min = float(self.dx_min)
        c.gamma = float(self.gamma)
        c.cfl = float(self.cfl)
        c.fixed_dt = float(self.fixed_dt or 0.0)
        c.recon = lookup(_lib.RECON, self.recon, "reconstruction_variable")
        c.stencil = lookup(_lib.STENCIL, self.stencil, "reconstruction_stencil")
        c.riemann = lookup(_lib.RIEMANN, self.riemann, "riemann_solver")
        c.signal_speed = lookup(_lib.SIGNAL, self.signal_speed, "signal_speed")
        c.integrator = lookup(_lib.INTEGRATOR, self.integrator, "integrator")
        c.convective_solver = lookup(_lib.CONVECTIVE_SOLVER, self.convective_solver, "convective_solver")
        c.flux_splitting = lookup(_lib.FLUX_SPLITTING, self.flux_splitting, "flux_splitting")
        c.frozen_state = lookup(_lib.FROZEN_STATE, self.frozen_state, "frozen_state")
        for k, f in enumerate(FACES):
            c.bc[k] = lookup(_lib.BC, self.bc[f], f"boundary condition type at {f}")
        c.viscous_flux = int(bool(self.is_viscous_flux))
        c.heat_flux = int(bool(self.is_heat_flux))
        c.viscous_heat_production = int(bool(self.is_viscous_heat_production))
        c.dynamic_viscosity = float(self.dynamic_viscosity)
        c.bulk_viscosity = float(self.bulk_viscosity)
        c.thermal_conductivity = self.thermal_conductivity_value() if self.is_heat_flux else 0.0
        c.gas_constant = float(self.gas_constant)
        c.interpolation_limiter = int(bool(self.is_interpolation_limiter))
        c.limit_velocity = int(bool(self.limit_velocity))
        if self.flux_limiter not in _lib.FLUX_LIMITER or self.flux_partition not in _lib.FLUX_PARTITION:
            raise NotImplementedError(f"positivity flux_limiter={self.flux_limiter!r} / flux_partition="
                                      f"{self.flux_partition!r} is not implemented on the B200 path")
        c.flux_limiter = _lib.FLUX_LIMITER[self.flux_limiter]
        c.flux_partition = _lib.FLUX_PARTITION[self.flux_partition]
        for k, f in enumerate(FACES):
            uvw = self.wall_velocity.get(f, (0.0, 0.0, 0.0))
            for q in range(3):
                c.wall_velocity[k][q] = float(uvw[q])
            vals = self.dirichlet.get(f, (1.0, 0.0, 0.0, 0.0, 1.0))
            for q in range(5):
                c.dirichlet[k][q] = float(vals[q])
        c.volume_force = int(bool(self.is_volume_force))
        c.no_convective_flux = int(not self.is_convective_flux)
        for q in range(3):
            c.gravity[q] = float(self.gravity[q])
        return c

    @property
    def shape(self):
        return (5,) + tuple(n + 2 * self.nh if n > 1 else 1 for n in self.cells)

    @property
    def rhs_shape(self):
        return (5,) + tuple(self.cells)

    @property
    def interior(self):
        return tuple(slice(self.nh, -self.nh) if n > 1 else slice(None) for n in self.cells)


def _ptr(t: Optional[torch.Tensor]):
    if t is None:
        return None
    assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous(), "fp64 contiguous CUDA tensor required"
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class BlockSolver:
    """Thin object wrapper over a jxf_handle."""

    def __init__(self, cfg: BlockConfig, device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise _lib.JxfError("jaxfluids_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.cfg = cfg
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.jxf_create(C.byref(cfg.to_c()), C.byref(self._h)))
        self.stages = self.lib.jxf_num_stages(self._h)
        self.active = tuple(i for i in range(3) if cfg.cells[i] > 1)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self.lib.jxf_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # -- buffers ----------------------------------------------------------
    def new_field(self, fill: Optional[float] = None) -> torch.Tensor:
        t = torch.empty(self.cfg.shape, dtype=torch.float64, device=self.device)
        if fill is not None:
            t.fill_(fill)
        return t

    def new_rhs(self) -> torch.Tensor:
        return torch.empty(self.cfg.rhs_shape, dtype=torch.float64, device=self.device)

    def new_scalars(self, n=1, value=0.0) -> torch.Tensor:
        return torch.full((n,), value, dtype=torch.float64, device=self.device)

    def new_red(self) -> torch.Tensor:
        return torch.tensor([0.0, float("inf"), float("inf")], dtype=torch.float64, device=self.device)

    # -- ABI calls (all enqueue-only on the current stream) ---------------
    def compute_rhs(self, prims: torch.Tensor, rhs: Optional[torch.Tensor] = None) -> torch.Tensor:
        rhs = self.new_rhs() if rhs is None else rhs
        _lib.check(self.lib.jxf_compute_rhs(self._h, _ptr(prims), _ptr(rhs), _stream()))
        return rhs

    def bind_timestep(self, dt_dev: Optional[torch.Tensor]):
        """Device scalar with the physical time step size that compute_rhs / sweep / sweep_range hand to the positivity
        flux limiter (the physical_timestep_size argument of SpaceSolver.compute_rhs, space_solver.py:151-164)."""
        self._dt_bound = dt_dev            # keep the tensor alive: the library stores the pointer only
        _lib.check(self.lib.jxf_bind_timestep(self._h, _ptr(dt_dev)))

    def sweep(self, axis: int, prims, rhs, accumulate: bool):
        _lib.check(self.lib.jxf_sweep(self._h, int(axis), _ptr(prims), _ptr(rhs), int(bool(accumulate)), _stream()))
        return rhs

    def stage(self, stage, prims_in, prims_out, cons_in, cons_n, cons_out, rhs, dt_dev, red_dev=None,
              reduce=False, fill_halo=True):
        _lib.check(self.lib.jxf_stage(self._h, int(stage), _ptr(prims_in), _ptr(prims_out), _ptr(cons_in),
                                      _ptr(cons_n), _ptr(cons_out), _ptr(rhs), _ptr(dt_dev), _ptr(red_dev),
                                      int(bool(reduce)), int(bool(fill_halo)), _stream()))

    def rhs_slab_elems(self, slab_planes: int) -> int:
        return int(self.lib.jxf_rhs_slab_elems(self._h, int(slab_planes)))

    def new_rhs_slabs(self, slab_planes: int) -> torch.Tensor:
        """Two slab-sized rhs accumulators, back to back (jxf_stage_inplace)."""
        return torch.empty(2 * self.rhs_slab_elems(slab_planes), dtype=torch.float64, device=self.device)

    def stage_inplace(self, stage, prims, cons_in, cons_n, cons_out, rhs_slabs, slab_planes, dt_dev, red_dev=None,
                      reduce=False, fill_halo=True):
        _lib.check(self.lib.jxf_stage_inplace(self._h, int(stage), _ptr(prims), _ptr(cons_in), _ptr(cons_n), _ptr(cons_out),
                                              _ptr(rhs_slabs), int(slab_planes), _ptr(dt_dev), _ptr(red_dev),
                                              int(bool(reduce)), int(bool(fill_halo)), _stream()))

    def sweep_range(self, axis: int, lo: int, hi: int, prims, rhs, accumulate: bool):
        _lib.check(self.lib.jxf_sweep_range(self._h, int(axis), int(lo), int(hi), _ptr(prims), _ptr(rhs),
                                            int(bool(accumulate)), _stream()))

    def stage_tail(self, stage, first_axis_index, prims_in, prims_out, cons_in, cons_n, cons_out, rhs, dt_dev,
                   red_dev=None, reduce=False, fill_halo=True):
        _lib.check(self.lib.jxf_stage_tail(self._h, int(stage), int(first_axis_index), _ptr(prims_in), _ptr(prims_out),
                                           _ptr(cons_in), _ptr(cons_n), _ptr(cons_out), _ptr(rhs), _ptr(dt_dev),
                                           _ptr(red_dev), int(bool(reduce)), int(bool(fill_halo)), _stream()))

    def step_fused(self, prims_a, prims_b, cons_a, cons_b, rhs, dt_dev, time_dev, red_dev, info_dev,
                   fill_halo=True) -> int:
        rc = self.lib.jxf_step_fused(self._h, _ptr(prims_a), _ptr(prims_b), _ptr(cons_a), _ptr(cons_b), _ptr(rhs),
                                     _ptr(dt_dev), _ptr(time_dev), _ptr(red_dev), _ptr(info_dev),
                                     int(bool(fill_halo)), _stream())
        if rc < 0:
            _lib.check(rc)
        return rc

    def integrate_stage(self, stage: int, cons, cons_n, rhs, dt: float, out=None):
        out = torch.empty_like(cons) if out is None else out
        _lib.check(self.lib.jxf_integrate_stage(self._h, int(stage), _ptr(cons), _ptr(cons_n), _ptr(rhs),
                                                C.c_double(float(dt)), _ptr(out), _stream()))
        return out

    def dissipative_sweep(self, axis: int, prims, rhs, accumulate: bool):
        _lib.check(self.lib.jxf_dissipative_sweep(self._h, int(axis), _ptr(prims), _ptr(rhs), int(bool(accumulate)),
                                                  _stream()))
        return rhs

    def halo_fill_edges(self, prims, cons):
        _lib.check(self.lib.jxf_halo_fill_edges(self._h, _ptr(prims), _ptr(cons), _stream()))

    def temperature(self, prims, out=None):
        out = torch.empty(tuple(prims.shape[1:]), dtype=torch.float64, device=prims.device) if out is None else out
        _lib.check(self.lib.jxf_temperature(self._h, _ptr(prims), _ptr(out), _stream()))
        return out

    def set_face_data(self, face: int, ops: int, data: Optional[torch.Tensor], mask: Optional[torch.Tensor] = None):
        """Per-face boundary data applied on top of the face's base rule (include/jxf_b200.h jxf_set_face_data);
        the caller keeps `data` (5, n1, n2) fp64 and `mask` (n1, n2) uint8 alive."""
        if data is not None:
            assert data.is_cuda and data.dtype == torch.float64 and data.is_contiguous()
        if mask is not None:
            assert mask.is_cuda and mask.dtype == torch.uint8 and mask.is_contiguous()
        _lib.check(self.lib.jxf_set_face_data(self._h, int(face), int(ops),
                                              None if data is None else C.c_void_p(data.data_ptr()),
                                              None if mask is None else C.c_void_p(mask.data_ptr())))

    # -- peer-memory halo exchange (include/jxf_b200.h) ------------------------
    def set_peer_halo(self, face: int, peer_prims_out: Optional[int], peer_cons_out: Optional[int]):
        """peer_*_out: peer-mapped device ADDRESSES (peer_import) of the neighbour's output buffers, or None"""
        _lib.check(self.lib.jxf_set_peer_halo(self._h, int(face),
                                              C.c_void_p(peer_prims_out) if peer_prims_out else None,
                                              C.c_void_p(peer_cons_out) if peer_cons_out else None))

    def peer_export(self, t: torch.Tensor):
        """-> (handle: 64 bytes, offset): names the device allocation holding t's first element for another process"""
        buf = C.create_string_buffer(64)
        off = C.c_int64(0)
        _lib.check(self.lib.jxf_peer_export(C.c_void_p(t.data_ptr()), buf, C.byref(off)))
        return bytes(buf.raw), int(off.value)

    def peer_import(self, handle: bytes, offset: int) -> int:
        """Map another process' allocation under the current device; -> the address of the exported element."""
        assert len(handle) == 64
        buf = C.create_string_buffer(handle, 64)
        out = C.c_void_p(0)
        _lib.check(self.lib.jxf_peer_import(buf, int(offset), C.byref(out)))
        return int(out.value)

    def peer_release(self):
        _lib.check(self.lib.jxf_peer_release())

    def peer_signal(self, slots, epoch: int):
        """slots: 6 device pointers (int) or None -- the neighbours' flag words this block writes"""
        arr = (C.c_void_p * 6)(*[C.c_void_p(p) if p else None for p in slots])
        _lib.check(self.lib.jxf_peer_signal(self._h, arr, int(epoch), _stream()))

    def peer_wait(self, flags: torch.Tensor, face_mask: int, epoch: int):
        _lib.check(self.lib.jxf_peer_wait(self._h, C.c_void_p(flags.data_ptr()), int(face_mask), int(epoch), _stream()))

    def halo_fill(self, prims, cons):
        _lib.check(self.lib.jxf_halo_fill(self._h, _ptr(prims), _ptr(cons), _stream()))

    def prims_from_cons(self, cons, prims):
        _lib.check(self.lib.jxf_prims_from_cons(self._h, _ptr(cons), _ptr(prims), _stream()))

    def cons_from_prims(self, prims, cons):
        _lib.check(self.lib.jxf_cons_from_prims(self._h, _ptr(prims), _ptr(cons), _stream()))

    def reduce(self, prims, red_dev):
        _lib.check(self.lib.jxf_reduce(self._h, _ptr(prims), _ptr(red_dev), _stream()))

    def reduce_reset(self, red_dev):
        _lib.check(self.lib.jxf_reduce_reset(self._h, _ptr(red_dev), _stream()))

    def finish_step(self, red_dev, dt_dev, time_dev=None, info_dev=None):
        _lib.check(self.lib.jxf_finish_step(self._h, _ptr(red_dev), _ptr(dt_dev), _ptr(time_dev), _ptr(info_dev),
                                            _stream()))

    PROFILE_KINDS = ("sweep_x", "sweep_y", "sweep_z", "sweep_x_epilogue", "sweep_y_epilogue", "sweep_z_epilogue",
                     "halo_fill", "other", "dissipative")

    def profile_enable(self, on: bool = True):
        _lib.check(self.lib.jxf_profile_enable(self._h, int(bool(on))))

    def profile_read(self, reset: bool = True):
        """-> {kind: (ms_sum, timed_launches, launches)}; synchronises on the recorded events."""
        n = len(self.PROFILE_KINDS)
        ms = (C.c_double * n)()
        timed = (C.c_int64 * n)()
        launches = (C.c_int64 * n)()
        _lib.check(self.lib.jxf_profile_read(self._h, ms, timed, launches, int(bool(reset))))
        return {k: (ms[i], timed[i], launches[i]) for i, k in enumerate(self.PROFILE_KINDS)}

    def face_slab_elems(self, face: int, ext_mask: int = 0, layers: Optional[int] = None) -> int:
        return int(self.lib.jxf_face_slab_elems_n(self._h, int(face), int(ext_mask), int(layers or self.cfg.nh)))

    def pack_face(self, face: int, prims, slab, ext_mask: int = 0, layers: Optional[int] = None):
        _lib.check(self.lib.jxf_pack_face_n(self._h, int(face), int(ext_mask), int(layers or self.cfg.nh), _ptr(prims),
                                            _ptr(slab), _stream()))

    def unpack_face(self, face: int, slab, prims, cons, ext_mask: int = 0, layers: Optional[int] = None):
        _lib.check(self.lib.jxf_unpack_face_n(self._h, int(face), int(ext_mask), int(layers or self.cfg.nh), _ptr(slab),
                                              _ptr(prims), _ptr(cons), _stream()))


class BlockState:
    """Ping-pong device state of one block + the no-sync single-block stepper."""

    def __init__(self, solver: BlockSolver, prims_with_halos: np.ndarray | torch.Tensor,
                 cons_with_halos: np.ndarray | torch.Tensor | None = None, dt: float | None = None, time: float = 0.0):
        s = self.solver = solver
        dev = s.device
        self.prims = [torch.as_tensor(prims_with_halos, dtype=torch.float64).to(dev).contiguous(), s.new_field(1.0)]
        self.cur = 0
        self.cons_a = s.new_field(1.0)
        self.cons_b = s.new_field(1.0)
        if cons_with_halos is not None:
            self.cons_a.copy_(torch.as_tensor(cons_with_halos, dtype=torch.float64))
        else:
            s.cons_from_prims(self.prims[0], self.cons_a)
        self.rhs = s.new_rhs() if (len(s.active) > 1 or s.cfg.is_dissipative) else None
        self.red = s.new_red()
        self.info = s.new_scalars(3)
        self.time = s.new_scalars(1, time)
        self.dt = s.new_scalars(1, 0.0)
        if s.cfg.flux_limiter:              # sweeps issued outside jxf_stage (overlap pieces, compute_rhs) read this dt
            s.bind_timestep(self.dt)
        if dt is None:
            self.update_dt()
        else:
            self.dt.fill_(dt)

    @property
    def primitives(self) -> torch.Tensor:
        return self.prims[self.cur]

    @property
    def conservatives(self) -> torch.Tensor:
        return self.cons_a

    def update_dt(self):
        """dt from the current primitives (time_control_initializer.py:34-78)."""
        s = self.solver
        s.reduce_reset(self.red)
        s.reduce(self.primitives, self.red)
        s.finish_step(self.red, self.dt, None, self.info)

    def step(self):
        """One full RK step, enqueue-only."""
        a, b = self.prims[self.cur], self.prims[self.cur ^ 1]
        where = self.solver.step_fused(a, b, self.cons_a, self.cons_b, self.rhs, self.dt, self.time, self.red, self.info)
        self.cur ^= where
