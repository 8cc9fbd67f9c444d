"""Containers crossing the API, with the reference's names and fields
(data_types/__init__.py:10-14, data_types/buffers.py:9-19,73-121,
data_types/information.py).  Arrays inside are torch CUDA fp64 tensors."""
from __future__ import annotations

from typing import Any, NamedTuple, Tuple


class MaterialFieldBuffers(NamedTuple):
    conservatives: Any = None
    primitives: Any = None
    temperature: Any = None


class LevelsetFieldBuffers(NamedTuple):
    levelset: Any = None
    volume_fraction: Any = None
    apertures: Any = None
    interface_velocity: Any = None
    interface_pressure: Any = None


class SolidFieldBuffers(NamedTuple):
    velocity: Any = None
    energy: Any = None
    temperature: Any = None


class SimulationBuffers(NamedTuple):
    material_fields: MaterialFieldBuffers = None
    levelset_fields: LevelsetFieldBuffers = None
    solid_fields: SolidFieldBuffers = None


class TimeControlVariables(NamedTuple):
    physical_simulation_time: float = None
    simulation_step: int = None
    physical_timestep_size: float = None
    fixed_time_step_size: float = None
    end_time: float = None
    end_step: int = None


class EulerIntegrationBuffers(NamedTuple):
    conservatives: Any = None
    levelset: Any = None
    solid_velocity: Any = None
    solid_energy: Any = None


class IntegrationBuffers(NamedTuple):
    euler_buffers: EulerIntegrationBuffers = None


class ForcingParameters(NamedTuple):
    mass_flow_controller_params: Any = None
    hit_ek_ref: Any = None


class PositivityCounter(NamedTuple):
    interpolation_limiter: int = None
    thinc_limiter: int = None
    flux_limiter: int = None
    acdi_limiter: int = None
    volume_fraction_limiter: int = None


class DiscretizationCounter(NamedTuple):
    acdi: int = None
    thinc: int = None


class PositivityStateInformation(NamedTuple):
    min_pressure: float = None
    min_density: float = None
    min_temperature: float = None
    min_alpharho: float = None
    min_alpha: float = None
    max_alpha: float = None
    positivity_counter: PositivityCounter = None
    discretization_counter: DiscretizationCounter = None
    levelset_fluid: Any = None
    levelset_solid: Any = None


class StepInformation(NamedTuple):
    positivity: Tuple[PositivityStateInformation] = ()
    levelset: Tuple = ()
    forcing_info: Any = None
    statistics: Any = None


class JaxFluidsBuffers(NamedTuple):
    simulation_buffers: SimulationBuffers
    time_control_variables: TimeControlVariables
    forcing_parameters: ForcingParameters
    step_information: StepInformation


class ControlFlowParameters(NamedTuple):
    perform_reinitialization: bool = False
    perform_compression: bool = False
    is_cumulative_statistics: bool = False
    is_logging_statistics: bool = False
    is_feed_foward: bool = False


class WallClockTimes(NamedTuple):
    step: float = 0.0
    step_per_cell: float = 0.0
    mean_step: float = 0.0
    mean_step_per_cell: float = 0.0
