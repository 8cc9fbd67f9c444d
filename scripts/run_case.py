#!/usr/bin/env python
"""Run a JAX-Fluids case file / numerical-setup file pair on the B200 path, the way the reference's
`examples/*/run.py` do (InputManager -> InitializationManager -> SimulationManager.simulate), and save the final
primitives (interior cells of this rank's block) to an .npz -- the HDF5 writer of the reference is out of scope
(no h5py in the image).

    python scripts/run_case.py path/to/case.json path/to/numerical_setup.json [--steps N] [--out result.npz]
    torchrun --nproc-per-node 4 scripts/run_case.py case.json numerical_setup.json      # decomposition from the case file
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("case")
    ap.add_argument("numerical_setup")
    ap.add_argument("--steps", type=int, default=None, help="override general/end_step (and lift end_time)")
    ap.add_argument("--out", default=None)
    ap.add_argument("--graph", action="store_true", help="replay the step from a CUDA graph (single block)")
    args = ap.parse_args()
    import numpy as np
    import torch
    import __graft_entry__ as entry
    entry.build()
    from jaxfluids_b200 import InputManager, InitializationManager, SimulationManager
    case = json.load(open(args.case))
    num = json.load(open(args.numerical_setup))
    if args.steps is not None:
        case["general"]["end_step"] = int(args.steps)
        case["general"]["end_time"] = 1e300
    im = InputManager(case, num)
    buffers = InitializationManager(im).initialization()
    sim = SimulationManager(im)
    if args.graph:
        sim.runtime.use_cuda_graph(True)
    torch.cuda.synchronize()
    t0 = time.time()
    sim.simulate(buffers)
    torch.cuda.synchronize()
    wall = time.time() - t0
    out = sim.final_buffers
    tcv = out.time_control_variables
    rt = sim.runtime
    prims = out.simulation_buffers.material_fields.primitives[(slice(None),) + rt.cfg.interior].cpu().numpy()
    cells = int(np.prod(im.domain_information.device_number_of_cells))
    print(json.dumps({"case": case["general"]["case_name"], "rank": rt.parallel.rank, "steps": tcv.simulation_step,
                      "time": tcv.physical_simulation_time, "dt": tcv.physical_timestep_size,
                      "min_density": out.step_information.positivity[-1].min_density,
                      "min_pressure": out.step_information.positivity[-1].min_pressure, "wall_s": wall,
                      "MCUPS_per_block": cells * max(tcv.simulation_step, 1) / max(wall, 1e-9) / 1e6}))
    if args.out:
        path = args.out if rt.parallel.world_size == 1 else f"{os.path.splitext(args.out)[0]}_rank{rt.parallel.rank}.npz"
        np.savez_compressed(path, primitives=prims, time=tcv.physical_simulation_time, step=tcv.simulation_step,
                            block_slices=np.array([[s.start, s.stop] for s in im.domain_information.block_slices(rt.parallel.rank)]))


if __name__ == "__main__":
    main()
