"""CPU: the log blocks keep the reference's layout (io_utils/logger.py:297-307, :324-361, :426-520, :603-615)."""
import logging

from jaxfluids_b200.data_types import (PositivityStateInformation, StepInformation, TimeControlVariables,
                                       WallClockTimes)
from jaxfluids_b200.logger import Logger


class _Capture(logging.Handler):
    def __init__(self):
        super().__init__()
        self.lines = []

    def emit(self, record):
        self.lines.append(record.getMessage())


def _logger(**kw):
    lg = Logger(logger_name="jaxfluids_b200_test", **kw)
    cap = _Capture()
    lg.logger.handlers[:] = [cap]
    lg.logger.propagate = False
    return lg, cap


def test_end_of_step_block_matches_the_reference_layout():
    lg, cap = _logger(level="INFO", frequency=2)
    tcv = TimeControlVariables(physical_simulation_time=0.0125, simulation_step=4, physical_timestep_size=3.5e-4,
                               fixed_time_step_size=False, end_time=1.0, end_step=100)
    info = StepInformation(positivity=(PositivityStateInformation(min_pressure=0.1, min_density=0.125),))
    lg.log_end_time_step(tcv, info, WallClockTimes(2.0e-3, 2.0e-9, 1.9e-3, 1.9e-9))
    star, blank = "*" + "-" * 78 + "*", f"{'*':<40}{'*':>40}"
    assert all(len(line) == 80 for line in cap.lines)
    assert cap.lines == [
        blank,
        f"*    {'TIME CONTROL':<74}*",
        f"*    {'CURRENT TIME                       = 1.25000e-02':<74}*",
        f"*    {'CURRENT DT                         = 3.50000e-04':<74}*",
        f"*    {'CURRENT STEP                       =      4':<74}*",
        f"*    {'WALL CLOCK TIMESTEP                = 2.00000e-03':<74}*",
        f"*    {'WALL CLOCK TIMESTEP CELL           = 2.00000e-09':<74}*",
        f"*    {'MEAN WALL CLOCK TIMESTEP CELL      = 1.90000e-09':<74}*",
        blank,
        blank,
        f"*    {'POSITIVITY STATE':<74}*",
        f"*    {'MIN DENSITY                        = 1.2500e-01':<74}*",
        f"*    {'MIN PRESSURE                       = 1.0000e-01':<74}*",
        blank,
        star,
    ]
    # the logging frequency of output/logging/frequency: odd steps are skipped
    cap.lines.clear()
    lg.log_end_time_step(tcv._replace(simulation_step=5), info, WallClockTimes())
    assert cap.lines == []


def test_level_none_and_non_root_ranks_are_silent():
    for kw in (dict(level="NONE"), dict(level="INFO", is_active=False)):
        lg, cap = _logger(**kw) if kw.get("level") != "NONE" and kw.get("is_active", True) else (Logger(
            logger_name="jaxfluids_b200_test_silent", **kw), None)
        handler = _Capture()
        lg.logger.addHandler(handler)
        lg.log_sim_finish(1.0)
        lg.hline()
        assert handler.lines == []
