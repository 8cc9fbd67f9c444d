"""Log blocks in the layout of the reference's logger (io_utils/logger.py): an 80-column star-framed stream with
TIME CONTROL and POSITIVITY STATE blocks per logged step, so that tooling which greps the reference's logs
("CURRENT TIME", "WALL CLOCK TIMESTEP CELL", "MIN DENSITY" ...) keeps working.

Covered: hline / nline / log_list (logger.py:297-307, :603-615), log_initial_time_step (:324-361), log_end_time_step for
the single-phase blocks (:426-470, :481-520), log_sim_start / log_sim_finish (:172-206; versions of this package).
Not covered: the ASCII banner, the setup dumps, level-set / forcing / turbulence-statistics blocks, *_TO_FILE levels."""
from __future__ import annotations

import logging
import os
import sys
import time
from typing import List


class Logger:
    def __init__(self, logger_name: str = "jaxfluids_b200", level: str = "INFO", frequency: int = 1,
                 is_positivity: bool = True, is_active: bool = True):
        self.level = level
        self.logging_frequency = max(int(frequency), 1)
        self.is_positivity = bool(is_positivity)
        self.is_active = bool(is_active) and level != "NONE"
        self.logger = logging.getLogger(logger_name)
        if self.is_active:
            if not self.logger.handlers:
                h = logging.StreamHandler(sys.stdout)
                h.setFormatter(logging.Formatter("%(message)s"))
                self.logger.addHandler(h)
            self.logger.setLevel(logging.DEBUG if "DEBUG" in level else logging.INFO)

    # -- primitives of the layout (logger.py:297-307, :603-618) -------------------
    def _emit(self, line: str) -> None:
        if self.is_active:
            self.logger.info(line)

    def hline(self) -> None:
        self._emit("*" + "-" * 78 + "*")

    def nline(self) -> None:
        self._emit(f"{'*':<40}{'*':>40}")

    def log_line(self, line: str) -> None:
        self._emit(f"*{line:<78}*")

    def log(self, line: str) -> None:
        self._emit(line)

    def log_list(self, input_list: List[str]) -> None:
        self.nline()
        for line in input_list:
            self._emit(f"*    {line:<74}*")
        self.nline()

    # -- blocks -----------------------------------------------------------------
    def log_sim_start(self, case_name: str = "", cells=None, blocks: int = 1) -> None:
        import torch
        from . import __version__ as version
        self.hline()
        self.nline()
        self._emit(f"*{'JAX-FLUIDS case files on the B200 path (jaxfluids_b200)':^78}*")
        self.nline()
        self.hline()
        self.nline()
        self._emit(f"*{'PYTHON Version: ' + sys.version.split()[0]:^78}*")
        self._emit(f"*{'TORCH Version: ' + torch.__version__:^78}*")
        self._emit(f"*{'jaxfluids_b200 Version: ' + str(version):^78}*")
        self._emit(f"*{'DATE & TIME: ' + time.strftime('%d/%m/%Y %H:%M:%S'):^78}*")
        self._emit(f"*{'PROCESS ID: ' + str(os.getpid()):^78}*")
        if case_name:
            self._emit(f"*{'CASE: ' + str(case_name):^78}*")
        if cells is not None:
            self._emit(f"*{'CELLS: ' + ' x '.join(str(int(c)) for c in cells) + f'  BLOCKS: {int(blocks)}':^78}*")
        self.nline()
        self.hline()

    def _time_control(self, tcv, time_reference: float) -> List[str]:
        return [
            "TIME CONTROL",
            f"CURRENT TIME                       = {tcv.physical_simulation_time * time_reference:4.5e}",
            f"CURRENT DT                         = {tcv.physical_timestep_size * time_reference:4.5e}",
            f"CURRENT STEP                       = {tcv.simulation_step:6d}",
        ]

    def _positivity(self, step_information) -> None:
        if not self.is_positivity or step_information is None or not step_information.positivity:
            return
        pos = step_information.positivity[-1]                      # logging.is_only_last_stage
        self.log_list(["POSITIVITY STATE",
                       f"MIN DENSITY                        = {pos.min_density:4.4e}",
                       f"MIN PRESSURE                       = {pos.min_pressure:4.4e}"])

    def log_initial_time_step(self, time_control_variables, step_information, time_reference: float = 1.0) -> None:
        self.log_list(self._time_control(time_control_variables, time_reference))
        self._positivity(step_information)
        self.hline()

    def log_end_time_step(self, time_control_variables, step_information, wall_clock_times,
                          time_reference: float = 1.0) -> None:
        if time_control_variables.simulation_step % self.logging_frequency != 0:
            return
        self.log_list(self._time_control(time_control_variables, time_reference) + [
            f"WALL CLOCK TIMESTEP                = {wall_clock_times.step:4.5e}",
            f"WALL CLOCK TIMESTEP CELL           = {wall_clock_times.step_per_cell:4.5e}",
            f"MEAN WALL CLOCK TIMESTEP CELL      = {wall_clock_times.mean_step_per_cell:4.5e}",
        ])
        self._positivity(step_information)
        self.hline()

    def log_sim_finish(self, end_time: float) -> None:
        self.hline()
        self.nline()
        self._emit(f"*{'SIMULATION FINISHED SUCCESSFULLY':^78}*")
        self._emit(f"*{f'SIMULATION TIME {end_time:.3e}s':^78}*")
        self.nline()
        self.hline()
