"""jaxfluids_b200 -- B200-native (sm_100a) drop-in for the single-phase convective
right-hand-side / SSP-RK path of JAX-Fluids (tumaer/JAXFLUIDS 0.2.1).

Public surface mirrors the reference (src/jaxfluids/__init__.py:52-63):

    from jaxfluids_b200 import InputManager, InitializationManager, SimulationManager
"""
__version__ = "0.1.0"

from .input_manager import InputManager  # noqa: E402
from .initialization_manager import InitializationManager  # noqa: E402
from .simulation_manager import SimulationManager  # noqa: E402

__all__ = ("InitializationManager", "InputManager", "SimulationManager")
