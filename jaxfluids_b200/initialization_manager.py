"""InitializationManager: builds the initial JaxFluidsBuffers of this rank's block.

Mirrors initialization/initialization_manager.py:142 ->
material_fields_initializer.py:590-690 (IC lambdas on the mesh grid) and
:425-497 / :148-210 (user-specified primitive array), then
time_control_initializer.py:34-78 (initial dt) and the initial positivity info
(initialization_manager.py:277-309).  The IC is evaluated on the host in NumPy
fp64 and uploaded once; prim->cons, halo fill and the dt reduction run in the
CUDA kernels.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .data_types import (ForcingParameters, JaxFluidsBuffers, LevelsetFieldBuffers, MaterialFieldBuffers,
                         PositivityStateInformation, SimulationBuffers, SolidFieldBuffers, StepInformation,
                         TimeControlVariables)
from .input_manager import InputManager
from .parallel import ParallelContext
from .runtime import BlockRuntime

EPS = float(np.finfo(np.float64).eps)      # config/precision.py:44-55


class InitializationManager:
    def __init__(self, input_manager: InputManager, parallel: Optional[ParallelContext] = None) -> None:
        self.input_manager = input_manager
        self.case_setup = input_manager.case_setup
        self.numerical_setup = input_manager.numerical_setup
        self.domain_information = input_manager.domain_information
        self.parallel = parallel if parallel is not None else ParallelContext.from_environment(self.domain_information)
        self.runtime = BlockRuntime.get(input_manager, self.parallel)

    # ------------------------------------------------------------------
    def _host_primitives_from_ic(self) -> np.ndarray:
        """(5, nx, ny, nz) block-interior primitives from the case file's initial condition
        (material_fields_initializer.py:626-643)."""
        di = self.domain_information
        mesh = di.compute_device_mesh_grid(self.parallel.rank, sparse=True)
        ic = self.case_setup.initial_condition_setup
        if "turbulent" in ic:
            return self._host_primitives_turbulent(ic["turbulent"])
        out = np.empty((5,) + tuple(di.device_number_of_cells), dtype=np.float64)
        shape = tuple(di.device_number_of_cells)
        for v, name in enumerate(("rho", "u", "v", "w", "p")):
            out[v] = np.broadcast_to(ic[name](*mesh), shape)
        return out

    def _host_primitives_turbulent(self, tb) -> np.ndarray:
        """initial_condition/turbulent/case = HIT (turbulence/initialization/turb_init_manager.py:44-82, hit.py:20-110):
        the generator works on the global grid (global FFTs); every rank evaluates it with the case file's seed and
        keeps its block."""
        from . import turbulence
        di = self.domain_information
        mat = self.case_setup.material_setup
        par = {k: v for k, v in tb.items() if k != "case"}
        glob = turbulence.initialize_hit(di.global_number_of_cells[0], mat.specific_heat_ratio,
                                         mat.specific_gas_constant, **par)
        return np.ascontiguousarray(glob[(slice(None),) + di.block_slices(self.parallel.rank)])

    def _host_primitives_from_user(self, user_prime_init) -> np.ndarray:
        """user_prime_init is the GLOBAL (5-3+dim, Nx, Ny, Nz) array (material_fields_initializer.py
        :425-497); inactive velocity components keep the eps fill of the buffer (:182-192)."""
        di = self.domain_information
        user = np.asarray(user_prime_init, dtype=np.float64)
        shape = (5 - 3 + di.dim,) + tuple(di.global_number_of_cells)
        assert user.shape == shape, (
            f"Given initial user primitive buffer has shape {user.shape} which is not consistent with the present "
            f"case setup file. The required shape is {shape}.")
        idx = [0] + [1 + i for i in di.active_axes_indices] + [4]
        sl = di.block_slices(self.parallel.rank)
        out = np.ones((5,) + tuple(di.device_number_of_cells), dtype=np.float64) * EPS
        out[idx] = user[(slice(None),) + sl]
        return out

    # ------------------------------------------------------------------
    def initialization(self, user_prime_init=None, user_time_init: Optional[float] = None,
                       user_levelset_init=None, user_solid_interface_velocity_init=None,
                       user_restart_file_path=None, ml_parameters=None, ml_callables=None) -> JaxFluidsBuffers:
        for name, v in (("user_levelset_init", user_levelset_init),
                        ("user_solid_interface_velocity_init", user_solid_interface_velocity_init),
                        ("user_restart_file_path", user_restart_file_path)):
            if v is not None:
                raise NotImplementedError(f"{name} is not implemented on the B200 path")
        if callable(user_prime_init):
            # extension of the reference's array argument (material_fields_initializer.py:425-497) for grids whose
            # GLOBAL array does not fit a host: fn(block_slices, out) fills `out`, the (5, bx, by, bz) interior VIEW of
            # this rank's device state buffer, with the block's primitives (all five rows; inactive velocities as given)
            sl = self.domain_information.block_slices(self.parallel.rank)
            host = lambda out: user_prime_init(sl, out)      # noqa: E731
        elif user_prime_init is not None:
            host = self._host_primitives_from_user(user_prime_init)
        else:
            host = self._host_primitives_from_ic()
        time0 = float(user_time_init) if isinstance(user_time_init, float) else 0.0

        rt = self.runtime
        prims, cons = rt.upload_initial_primitives(host)        # eps-filled buffers, prim->cons, halo update
        dt, min_rho, min_p = rt.initial_time_step_and_positivity(prims)

        gs = self.case_setup.general_setup
        fixed = self.numerical_setup.conservatives.time_integration.fixed_timestep
        tcv = TimeControlVariables(
            physical_simulation_time=time0, simulation_step=0, physical_timestep_size=dt,
            fixed_time_step_size=fixed, end_time=gs.end_time, end_step=gs.end_step)
        # equation_information.is_compute_temperature: the temperature buffer exists with the viscous / heat flux
        temperature = rt.temperature(prims)
        material_fields = MaterialFieldBuffers(conservatives=cons, primitives=prims, temperature=temperature)
        sim = SimulationBuffers(material_fields, LevelsetFieldBuffers(), SolidFieldBuffers())
        step_info = StepInformation(positivity=(PositivityStateInformation(min_pressure=min_p, min_density=min_rho),))
        return JaxFluidsBuffers(sim, tcv, ForcingParameters(), step_info)
