"""CPU: the host-side step loop of SimulationManager.simulate (sequencing, time control, step information, logging) with a
stub runtime in place of the CUDA one -- catches Python-level errors of the driver glue (simulation_manager.py:186-465 of
the reference) without a GPU.  The arithmetic is NOT exercised here (tests -m gpu do that)."""
import json
import logging
from types import SimpleNamespace

import numpy as np
import torch

from jaxfluids_b200 import InputManager, SimulationManager
from jaxfluids_b200.data_types import (ForcingParameters, JaxFluidsBuffers, LevelsetFieldBuffers, MaterialFieldBuffers,
                                       PositivityStateInformation, SimulationBuffers, SolidFieldBuffers, StepInformation,
                                       TimeControlVariables)
from jaxfluids_b200.parallel import ParallelContext
from jaxfluids_b200.runtime import BlockRuntime
from tests import helpers as H


class StubRuntime:
    """What SimulationManager touches of BlockRuntime; advances time by the current dt and halves dt every step."""

    def __init__(self, im):
        self.parallel = ParallelContext(im.domain_information)
        self.stages = 3
        shape = (5,) + tuple(n + 10 if n > 1 else 1 for n in im.domain_information.device_number_of_cells)
        self.primitives = torch.ones(shape, dtype=torch.float64)
        self.conservatives = torch.ones(shape, dtype=torch.float64)
        self.cfg = SimpleNamespace(is_dissipative=False)
        self.t, self.dt, self.steps = 0.0, 0.0, 0

    def adopt(self, p, c):
        pass

    def set_time_control(self, t, dt):
        self.t, self.dt = float(t), float(dt)

    def step(self):
        self.t += self.dt
        self.dt *= 0.5
        self.steps += 1

    def read_step_scalars(self, complete_halos=False):
        return self.t, self.dt, 3.0, 0.125, 0.1

    def complete_halos(self):
        pass

    def temperature(self, prims):
        return None


def test_simulate_runs_the_step_loop_and_logs_in_the_reference_layout(monkeypatch, capsys):
    g, case, num = H.load_golden("sod200_char_hllc_rk3")
    case, num = json.loads(json.dumps(case)), json.loads(json.dumps(num))
    case["general"]["end_step"] = 4
    case["general"]["end_time"] = 1e300
    num.setdefault("output", {})["logging"] = {"level": "INFO", "frequency": 2}
    im = InputManager(case, num)
    stub = StubRuntime(im)
    monkeypatch.setattr(BlockRuntime, "get", classmethod(lambda cls, input_manager, parallel=None: stub))
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    logging.getLogger("jaxfluids_b200").handlers[:] = []
    sim = SimulationManager(im)
    mf = MaterialFieldBuffers(conservatives=stub.conservatives, primitives=stub.primitives, temperature=None)
    tcv = TimeControlVariables(physical_simulation_time=0.0, simulation_step=0, physical_timestep_size=0.01,
                               fixed_time_step_size=False, end_time=im.case_setup.general_setup.end_time,
                               end_step=im.case_setup.general_setup.end_step)
    info = StepInformation(positivity=(PositivityStateInformation(min_pressure=0.1, min_density=0.125),))
    buffers = JaxFluidsBuffers(SimulationBuffers(mf, LevelsetFieldBuffers(), SolidFieldBuffers()), tcv, ForcingParameters(),
                               info)
    assert sim.simulate(buffers) == 0
    out = sim.final_buffers.time_control_variables
    assert stub.steps == 4 and out.simulation_step == 4
    assert abs(out.physical_simulation_time - 0.01 * (1 + 0.5 + 0.25 + 0.125)) < 1e-15
    assert out.physical_timestep_size == 0.01 / 16
    assert sim.final_buffers.step_information.positivity[-1].min_density == 0.125
    text = capsys.readouterr().out
    lines = [ln for ln in text.splitlines() if ln]
    assert all(len(ln) == 80 and ln[0] == "*" and ln[-1] == "*" for ln in lines)
    assert sum("CURRENT STEP" in ln for ln in lines) == 3            # the initial block + steps 2 and 4 (frequency 2)
    assert any("SIMULATION FINISHED SUCCESSFULLY" in ln for ln in lines)
    assert any("MIN PRESSURE                       = 1.0000e-01" in ln for ln in lines)
    logging.getLogger("jaxfluids_b200").handlers[:] = []
