"""Callback base class with the reference's hook names (callbacks/base_callback.py:21-240).

Hooks the B200 path calls (simulation_manager.py:312-407, :565-665, :778, :1018 in the reference):
`on_simulation_start`, `on_simulation_end`, `before_step_start`, `after_step_end`, `on_step_start`, `on_step_end`
with `(jxf_buffers, callback_dict)`, and `on_stage_start` / `on_stage_end` with the stage's conservatives and
primitives.  `after_compute_rhs` would need the right-hand side as a separate array between the last sweep and the stage
update, which the fused stage kernel never materialises: a callback that overrides it is rejected at construction.
Every hook defaults to the identity, so unknown / empty callbacks cost nothing.
"""
from __future__ import annotations

from typing import Dict, Tuple


class Callback:
    def init_callback(self, sim_manager) -> None:
        self.sim_manager = sim_manager
        self.domain_information = sim_manager.domain_information
        self.equation_information = sim_manager.equation_information

    def on_simulation_start(self, jxf_buffers, callback_dict: Dict, **kwargs) -> Tuple[object, Dict]:
        return jxf_buffers, callback_dict

    def on_simulation_end(self, jxf_buffers, callback_dict: Dict, **kwargs) -> Tuple[object, Dict]:
        return jxf_buffers, callback_dict

    def on_step_start(self, jxf_buffers, callback_dict: Dict, **kwargs) -> Tuple[object, Dict]:
        return jxf_buffers, callback_dict

    def on_step_end(self, jxf_buffers, callback_dict: Dict, **kwargs) -> Tuple[object, Dict]:
        return jxf_buffers, callback_dict

    def before_step_start(self, jxf_buffers, callback_dict: Dict, **kwargs) -> Tuple[object, Dict]:
        return jxf_buffers, callback_dict

    def after_step_end(self, jxf_buffers, callback_dict: Dict, **kwargs) -> Tuple[object, Dict]:
        return jxf_buffers, callback_dict

    def on_stage_start(self, conservatives, primitives, physical_timestep_size=None, physical_simulation_time=None,
                       levelset=None, volume_fraction=None, apertures=None, forcing_buffers=None, ml_setup=None,
                       **kwargs):
        return conservatives, primitives

    def on_stage_end(self, conservatives, primitives, physical_timestep_size=None, physical_simulation_time=None,
                     levelset=None, volume_fraction=None, apertures=None, forcing_buffers=None, ml_setup=None,
                     **kwargs):
        return conservatives, primitives

    def on_rhs_axis(self) -> None:
        return None

    def after_compute_rhs(self, rhs_buffers, material_fields=None, levelset_fields=None, solid_fields=None,
                          forcing_buffers=None, ml_setup=None, **kwargs):
        return rhs_buffers


HOOKS = ("on_simulation_start", "on_simulation_end", "on_step_start", "on_step_end", "before_step_start",
         "after_step_end", "on_stage_start", "on_stage_end")


def overrides(cb, name: str) -> bool:
    """True when `cb` defines hook `name` itself (not the identity inherited from Callback / absent)."""
    fn = getattr(type(cb), name, None)
    return fn is not None and fn is not getattr(Callback, name, None)
