#!/usr/bin/env python
"""MCUPS of the other BASELINE configs on one GPU (parity-test cases, not the bench line):
Sod 1000, 2-D Riemann 1024^2, TGV 256^3.  Prints one JSON line per config."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402
import bench  # noqa: E402

entry.build()
from jaxfluids_b200 import InputManager, InitializationManager, SimulationManager  # noqa: E402

NUM = bench.tgv_case(32, (1, 1, 1), 10)[1]


def sod(n):
    return {"general": {"case_name": "sod", "end_step": 10 ** 9, "save_path": "./results"},
            "domain": {"x": {"cells": n, "range": [0.0, 1.0]}, "y": {"cells": 1, "range": [0.0, 1.0]},
                       "z": {"cells": 1, "range": [0.0, 1.0]}},
            "boundary_conditions": {"east": {"type": "ZEROGRADIENT"}, "west": {"type": "ZEROGRADIENT"},
                                    "north": {"type": "INACTIVE"}, "south": {"type": "INACTIVE"},
                                    "top": {"type": "INACTIVE"}, "bottom": {"type": "INACTIVE"}},
            "initial_condition": {"rho": "lambda x: 1.0*(x <= 0.5) + 0.125*(x > 0.5)", "u": 0.0, "v": 0.0, "w": 0.0,
                                  "p": "lambda x: 1.0*(x <= 0.5) + 0.1*(x > 0.5)"},
            "material_properties": {"equation_of_state": {"model": "IdealGas", "specific_heat_ratio": 1.4,
                                                          "specific_gas_constant": 1.0}}}


def riemann2d(n):
    q = lambda a, b, c, d: (f"lambda x, y: ((x >= 0.5) & (y >= 0.5)) * {a} + ((x < 0.5) & (y >= 0.5)) * {b} + "
                            f"((x < 0.5) & (y < 0.5)) * {c} + ((x >= 0.5) & (y < 0.5)) * {d}")
    c = sod(n)
    c["general"]["case_name"] = "riemann2D"
    c["domain"]["y"]["cells"] = n
    for f in ("north", "south"):
        c["boundary_conditions"][f] = {"type": "ZEROGRADIENT"}
    c["initial_condition"] = {"rho": q(1.5, 0.5323, 0.138, 0.5323), "u": q(0.0, 1.206, 1.206, 0.0),
                              "v": q(0.0, 0.0, 1.206, 1.206), "w": 0.0, "p": q(1.5, 0.3, 0.029, 0.3)}
    return c


def run(name, case, steps, warmup, graph=False):
    im = InputManager(case, NUM)
    buf = InitializationManager(im).initialization()
    sim = SimulationManager(im)
    rt = sim.runtime
    if graph:
        rt.use_cuda_graph(True)
        name += " [CUDA graph]"
    tcv = buf.time_control_variables
    rt.set_time_control(tcv.physical_simulation_time, tcv.physical_timestep_size)
    for _ in range(warmup):
        rt.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        rt.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    cells = im.domain_information.cells_per_device
    t, dt, _, mr, mp = rt.read_step_scalars()
    print(json.dumps({"config": name, "cells": cells, "ms_per_step": ms, "MCUPS": cells / ms / 1e3, "time": t, "dt": dt,
                      "min_density": mr, "min_pressure": mp}), flush=True)


if __name__ == "__main__":
    run("Sod 1000 cells", sod(1000), 200, 20)
    run("Sod 1000 cells", sod(1000), 200, 20, graph=True)
    run("2-D Riemann 1024^2", riemann2d(1024), 50, 5)
    run("2-D Riemann 1024^2", riemann2d(1024), 50, 5, graph=True)
    run("TGV 256^3", bench.tgv_case(256, (1, 1, 1), 10 ** 9)[0], 10, 3)
