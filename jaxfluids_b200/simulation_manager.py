"""SimulationManager and the component objects of the hot path, with the reference's
names, signatures and return containers (simulation_manager.py:186-1246,
solvers/space_solver.py:151-169, time_integration/time_integrator.py:108-126,
halos/halo_manager.py:146, equation_manager.py:93-171, time_step_size.py:15).

Every array method enqueues sm_100a kernels through the C ABI; nothing here
computes fields on the host.  The `simulate` loop advances with the fused
per-step driver and syncs with the device once per step for the loop condition and the
log line, as the reference's host loop does (simulation_manager.py:325-397).
"""
from __future__ import annotations

import contextlib
import time as _time
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from .data_types import (ControlFlowParameters, DiscretizationCounter, EulerIntegrationBuffers, ForcingParameters,
                         IntegrationBuffers, JaxFluidsBuffers, LevelsetFieldBuffers, MaterialFieldBuffers,
                         PositivityCounter, PositivityStateInformation, SimulationBuffers, SolidFieldBuffers,
                         StepInformation, TimeControlVariables, WallClockTimes)
from . import callbacks as _cb
from .input_manager import InputManager
from .logger import Logger
from .parallel import ParallelContext
from .runtime import BlockRuntime


class EquationManager:
    """equation_manager.py:13 (SINGLE-PHASE branches)."""

    def __init__(self, runtime: BlockRuntime, equation_information):
        self._rt = runtime
        self.equation_information = equation_information

    def get_conservatives_from_primitives(self, primitives: torch.Tensor) -> torch.Tensor:
        out = torch.empty_like(primitives)
        self._rt.solver.cons_from_prims(primitives.contiguous(), out)
        return out

    def get_primitives_from_conservatives(self, conservatives: torch.Tensor) -> torch.Tensor:
        out = torch.empty_like(conservatives)
        self._rt.solver.prims_from_cons(conservatives.contiguous(), out)
        return out


class SpaceSolver:
    """solvers/space_solver.py: convective single-phase right-hand side."""

    def __init__(self, runtime: BlockRuntime):
        self._rt = runtime

    def compute_rhs(self, conservatives, primitives, temperature=None, physical_simulation_time=0.0,
                    physical_timestep_size=0.0, levelset=None, volume_fraction=None, apertures=None,
                    interface_velocity=None, interface_pressure=None, solid_velocity=None, solid_temperature=None,
                    interface_cells=None, forcing_buffers=None, ml_setup=None, is_feed_forward=False
                    ) -> Tuple[IntegrationBuffers, PositivityCounter, DiscretizationCounter]:
        """space_solver.py:151-453.  `conservatives` is not read: on this path the reference uses it for the volume forces
        and the positivity flux limiter only (:378-384, :532-543), where it is U(primitives) -- the kernels form those
        values from `primitives` themselves (same arithmetic as equation_manager.get_conservatives_from_primitives).  A
        caller that passes conservatives inconsistent with its primitives gets the primitives' answer."""
        with self._timestep(physical_timestep_size):
            rhs = self._rt.solver.compute_rhs(primitives)
        return (IntegrationBuffers(EulerIntegrationBuffers(rhs, None, None, None)), PositivityCounter(),
                DiscretizationCounter())

    def compute_rhs_xi(self, conservatives, primitives, temperature, axis, physical_simulation_time=0.0,
                       physical_timestep_size=0.0, *args, **kwargs):
        rhs = self._rt.solver.new_rhs()
        with self._timestep(physical_timestep_size):
            self._rt.solver.sweep(axis, primitives, rhs, accumulate=False)
        return EulerIntegrationBuffers(rhs, None, None, None), PositivityCounter(), DiscretizationCounter()

    @contextlib.contextmanager
    def _timestep(self, physical_timestep_size):
        """The positivity flux limiter scales with the physical time step size (space_solver.py:532-543): bind the
        caller's value for the duration of the call, then the stepper's own device scalar again."""
        solver = self._rt.solver
        if not solver.cfg.flux_limiter:
            yield
            return
        if torch.is_tensor(physical_timestep_size):
            dt = physical_timestep_size.to(device=solver.device, dtype=torch.float64).reshape(1).contiguous()
        else:
            dt = solver.new_scalars(1, float(physical_timestep_size))
        previous = getattr(solver, "_dt_bound", None)
        solver.bind_timestep(dt)
        try:
            yield
        finally:
            solver.bind_timestep(previous)     # (the scratch scalar is freed stream-ordered, after the sweeps read it)


class TimeIntegrator:
    """time_integration/time_integrator.py + RK3.py / RK2.py / euler.py / RK2_LS4.py tables."""
    TABLES = {
        "EULER": (1, (1.0,), (1.0,), ()),
        "RK2": (2, (1.0, 0.5), (1.0, 1.0), ((0.5, 0.5),)),
        "RK3": (3, (1.0, 0.25, 2.0 / 3.0), (1.0, 0.5, 1.0), ((0.25, 0.75), (2.0 / 3.0, 1.0 / 3.0))),
        "RK2_LS4": (4, (0.11, 0.2766, 0.5, 1.0), (0.11, 0.2766, 0.5, 1.0), ((0.0, 1.0), (0.0, 1.0), (0.0, 1.0))),
    }

    def __init__(self, runtime: BlockRuntime, name: str):
        self._rt = runtime
        self.name = name
        self.no_stages, self.timestep_multiplier, self.timestep_increment_factor, self.conservatives_multiplier = \
            self.TABLES[name]

    def perform_stage_integration(self, integration_buffers: IntegrationBuffers, rhs_buffers: IntegrationBuffers,
                                  initial_stage_buffers: Optional[IntegrationBuffers], physical_timestep_size,
                                  stage: int, equation_information=None) -> IntegrationBuffers:
        """time_integrator.py:108-227: stage > 0: U <- a U + b U^n on the whole buffer (RK3.py:49-50), then
        interior U += (dt m_s) rhs (time_integrator.py:57).  Returns NEW buffers like the reference (the
        production step fuses this into the last sweep kernel; this is the stand-alone entry)."""
        cons = integration_buffers.euler_buffers.conservatives
        cons_n = initial_stage_buffers.euler_buffers.conservatives if stage > 0 else None
        rhs = rhs_buffers.euler_buffers.conservatives
        out = self._rt.solver.integrate_stage(stage, cons.contiguous(), None if cons_n is None else cons_n.contiguous(),
                                              rhs.contiguous(), float(physical_timestep_size))
        return IntegrationBuffers(EulerIntegrationBuffers(out, None, None, None))


class HaloManager:
    """halos/halo_manager.py:146-234."""

    def __init__(self, runtime: BlockRuntime):
        self._rt = runtime
        self.fill_edge_halos_material = runtime.cfg.is_dissipative      # halo_manager.py:119-129
        self.fill_vertex_halos_material = False                         # :131-143: level-set models only

    def perform_halo_update_material(self, primitives, physical_simulation_time=0.0, fill_edge_halos=False,
                                     fill_vertex_halos=False, conservatives=None, fill_face_halos=True,
                                     ml_setup=None):
        assert not fill_vertex_halos, "vertex halos are not on this path (level-set models only)"
        cons = conservatives if conservatives is not None else torch.empty_like(primitives)
        self._rt.halo_update(primitives, cons)                 # faces; + edges when the dissipative fluxes are on
        if fill_edge_halos and not self._rt.cfg.is_dissipative:
            self._rt.solver.halo_fill_edges(primitives, cons)
        return (primitives, cons) if conservatives is not None else primitives

    def perform_outer_halo_update_temperature(self, temperature, physical_simulation_time=0.0):
        """halo_manager.py:236-253.  For PERIODIC / SYMMETRY / ZEROGRADIENT faces the temperature halos equal the
        temperature of the halo primitives, which `get_temperature` on the halo'd buffer already produced."""
        return temperature


def compute_time_step_size(primitives, runtime: BlockRuntime) -> float:
    """time_integration/time_step_size.py:15-157 on the device; returns the host float."""
    dt, _, _ = runtime.initial_time_step_and_positivity(primitives)
    return dt


class SimulationManager:
    def __init__(self, input_manager: InputManager, callbacks=None, parallel: Optional[ParallelContext] = None) -> None:
        # simulation_manager.py:179-184: one callback or a list; every hook of the reference's Callback is honoured
        # except after_compute_rhs (the fused stage kernel never materialises the rhs between sweep and update)
        if callbacks is not None and not isinstance(callbacks, (list, tuple)):
            callbacks = [callbacks]
        self.callbacks = list(callbacks or [])
        for cb in self.callbacks:
            if _cb.overrides(cb, "after_compute_rhs"):
                raise NotImplementedError("callback hook 'after_compute_rhs' is not available on the B200 path: the stage "
                                          "kernel fuses the last sweep with the stage update")
        self._stage_hooks = any(_cb.overrides(cb, h) for cb in self.callbacks for h in ("on_stage_start", "on_stage_end"))
        self.input_manager = input_manager
        self.case_setup = input_manager.case_setup
        self.numerical_setup = input_manager.numerical_setup
        self.domain_information = input_manager.domain_information
        self.equation_information = input_manager.equation_information
        self.runtime = BlockRuntime.get(input_manager, parallel)
        self.parallel = self.runtime.parallel
        rt = self.runtime
        self.equation_manager = EquationManager(rt, self.equation_information)
        self.space_solver = SpaceSolver(rt)
        self.time_integrator = TimeIntegrator(rt, self.numerical_setup.conservatives.time_integration.integrator)
        self.halo_manager = HaloManager(rt)
        log = self.numerical_setup.output.logging
        # the reference's block layout (io_utils/logger.py); rank 0 only, like the reference's is_multihost guard
        self.logger = Logger(level=log.level, frequency=log.frequency, is_positivity=log.is_positivity,
                             is_active=self.parallel.rank == 0)
        self.wall_clock_times = WallClockTimes()
        for cb in self.callbacks:
            if hasattr(cb, "init_callback"):
                cb.init_callback(sim_manager=self)

    def _callback(self, hook_name: str, jxf_buffers=None, callback_dict=None, conservatives=None, primitives=None,
                  **kwargs):
        """simulation_manager.py:1079-1160: run `hook_name` of every callback (missing hooks are the identity)."""
        if hook_name in ("on_stage_start", "on_stage_end"):
            for cb in self.callbacks:
                fn = getattr(cb, hook_name, None)
                if fn is not None:
                    conservatives, primitives = fn(conservatives=conservatives, primitives=primitives, **kwargs)
            return conservatives, primitives
        for cb in self.callbacks:
            fn = getattr(cb, hook_name, None)
            if fn is not None:
                jxf_buffers, callback_dict = fn(jxf_buffers=jxf_buffers, callback_dict=callback_dict, **kwargs)
        return jxf_buffers, callback_dict

    # ------------------------------------------------------------------
    def simulate(self, jxf_buffers: JaxFluidsBuffers, ml_parameters=None, ml_callables=None) -> int:
        """simulation_manager.py:186-295 (no h5 output on this path)."""
        self.logger.log_sim_start(self.case_setup.general_setup.case_name,
                                  self.domain_information.global_number_of_cells, self.parallel.world_size)
        self.logger.log_initial_time_step(jxf_buffers.time_control_variables, jxf_buffers.step_information)
        t0 = _time.time()
        self.final_buffers = self.advance(jxf_buffers, ml_parameters, ml_callables)
        self.logger.log_sim_finish(_time.time() - t0)
        return 0

    def advance(self, jxf_buffers: JaxFluidsBuffers, ml_parameters=None, ml_callables=None) -> JaxFluidsBuffers:
        """simulation_manager.py:297-426: host while-loop over steps."""
        tcv = jxf_buffers.time_control_variables
        t, step = tcv.physical_simulation_time, tcv.simulation_step
        cells = self.domain_information.cells_per_device
        n_timed, mean = 0, 0.0
        callback_dict: Dict = {}
        jxf_buffers, callback_dict = self._callback("on_simulation_start", jxf_buffers, callback_dict)
        while t < tcv.end_time and step < tcv.end_step:
            torch.cuda.synchronize()
            t0 = _time.time()
            cfp = self.compute_control_flow_params(tcv, jxf_buffers.step_information)
            jxf_buffers, callback_dict = self._callback("before_step_start", jxf_buffers, callback_dict)
            jxf_buffers, callback_dict_step = self._do_integration_step(jxf_buffers, cfp, ml_parameters, ml_callables,
                                                                        complete_halos=bool(self.callbacks))
            jxf_buffers, callback_dict = self._callback("after_step_end", jxf_buffers, callback_dict,
                                                        callback_dict_step=callback_dict_step)
            tcv = jxf_buffers.time_control_variables
            t, step = tcv.physical_simulation_time, tcv.simulation_step
            wall = _time.time() - t0
            if step > 10:                                   # simulation_manager.py:428-465: skip warm-up steps
                n_timed += 1
                mean += (wall - mean) / n_timed
            self.wall_clock_times = WallClockTimes(wall, wall / cells, mean, mean / cells)
            self.logger.log_end_time_step(tcv, jxf_buffers.step_information, self.wall_clock_times)
        self.runtime.complete_halos()
        jxf_buffers, callback_dict = self._callback("on_simulation_end", jxf_buffers, callback_dict)
        return jxf_buffers

    def compute_control_flow_params(self, time_control_variables, step_information) -> ControlFlowParameters:
        return ControlFlowParameters()

    # ------------------------------------------------------------------
    def do_integration_step(self, jxf_buffers: JaxFluidsBuffers, control_flow_params=None, ml_parameters=None,
                            ml_callables=None) -> Tuple[JaxFluidsBuffers, Dict]:
        """simulation_manager.py:1178-1246 -> _do_integration_step :536-668.

        The returned JaxFluidsBuffers hold VIEWS of the runtime's device buffers (no copy): the next step overwrites
        them, where the reference returns fresh immutable arrays.  Clone what must outlive the next call."""
        return self._do_integration_step(jxf_buffers, control_flow_params, ml_parameters, ml_callables)

    def _do_integration_step(self, jxf_buffers, control_flow_params=None, ml_parameters=None, ml_callables=None,
                             complete_halos: bool = True):
        rt = self.runtime
        callback_dict: Dict = {}
        jxf_buffers, callback_dict = self._callback("on_step_start", jxf_buffers, callback_dict)
        mf = jxf_buffers.simulation_buffers.material_fields
        tcv = jxf_buffers.time_control_variables
        rt.adopt(mf.primitives, mf.conservatives)
        rt.set_time_control(tcv.physical_simulation_time, tcv.physical_timestep_size)
        if self._stage_hooks:
            self._step_with_stage_hooks(tcv)
        else:
            rt.step()
        # multi-block runs ship only the layers the stencils read between stages; complete the rest before the
        # buffers go back to the caller (advance() defers this to the end of its loop when no callback looks)
        t, dt_next, _, min_rho, min_p = rt.read_step_scalars(complete_halos=complete_halos)
        tcv = tcv._replace(physical_simulation_time=t, simulation_step=tcv.simulation_step + 1,
                           physical_timestep_size=dt_next)
        material_fields = MaterialFieldBuffers(rt.conservatives, rt.primitives, rt.temperature(rt.primitives))
        sim = SimulationBuffers(material_fields, jxf_buffers.simulation_buffers.levelset_fields,
                                jxf_buffers.simulation_buffers.solid_fields)
        info = StepInformation(positivity=(PositivityStateInformation(min_pressure=min_p, min_density=min_rho),))
        out = JaxFluidsBuffers(sim, tcv, jxf_buffers.forcing_parameters, info)
        return self._callback("on_step_end", out, callback_dict)

    def _step_with_stage_hooks(self, tcv):
        """The step stage by stage with on_stage_start / on_stage_end around every stage (simulation_manager.py:778,
        :1018).  The hooks see the stage's conservatives / primitives; what they return becomes the state."""
        rt = self.runtime
        kw = dict(physical_timestep_size=tcv.physical_timestep_size, physical_simulation_time=tcv.physical_simulation_time)
        def hook(name, cons, prims):
            c, p = self._callback(name, conservatives=cons, primitives=prims, **kw)
            if c is not cons:
                cons.copy_(c)
            if p is not prims:
                prims.copy_(p)
        for k in range(rt.stages):
            last = k == rt.stages - 1
            rt.finish_pending()
            hook("on_stage_start", rt.cons[0] if k == 0 else rt.cons[1], rt.primitives)
            rt.stage(k, reduce=last)
            rt.finish_pending()
            hook("on_stage_end", rt.cons[0] if last else rt.cons[1], rt.primitives)
        rt._allreduce_red()
        rt.solver.finish_step(rt.red, rt.dt, rt.time, rt.info)

    def do_runge_kutta_stages(self, material_fields: MaterialFieldBuffers, time_control_variables: TimeControlVariables,
                              levelset_fields=None, solid_fields=None, forcing_buffers=None,
                              control_flow_params=None, ml_setup=None):
        """simulation_manager.py:670-1077, single-phase branch: all stages, t += dt, step += 1 (no new dt)."""
        rt = self.runtime
        rt.adopt(material_fields.primitives, material_fields.conservatives)
        rt.set_time_control(time_control_variables.physical_simulation_time,
                            time_control_variables.physical_timestep_size)
        rt.solver.reduce_reset(rt.red)
        for k in range(rt.stages):
            rt.stage(k, reduce=(k == rt.stages - 1))
        rt._allreduce_red()
        rt.complete_halos()
        red = rt.red.cpu().numpy()
        tcv = time_control_variables._replace(
            physical_simulation_time=time_control_variables.physical_simulation_time +
            time_control_variables.physical_timestep_size,
            simulation_step=time_control_variables.simulation_step + 1)
        info = StepInformation(positivity=(PositivityStateInformation(min_pressure=float(red[2]),
                                                                      min_density=float(red[1])),))
        return (MaterialFieldBuffers(rt.conservatives, rt.primitives, rt.temperature(rt.primitives)), tcv,
                levelset_fields, solid_fields, info)
