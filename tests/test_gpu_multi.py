"""GPU, multi-rank: block decomposition with NCCL halo exchange against the single-block oracle.
Launched as a subprocess with torchrun so that a plain `pytest -m gpu` covers it; skipped when fewer
than 2 GPUs are visible."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu
ROOT = H.ROOT


def _ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


WORKER = r'''
import json, os, sys, signal
signal.alarm(300)          # watchdog: a rank stuck in an exchange dies by itself instead of holding the GPU
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["JXF_ROOT"])
from jaxfluids_b200 import InputManager, InitializationManager, SimulationManager
from oracle import port
from tests import helpers as H

split = tuple(int(v) for v in os.environ["JXF_SPLIT"].split(","))
bc = os.environ["JXF_BC"]
nsteps = int(os.environ["JXF_STEPS"])
cells = tuple(int(v) for v in os.environ["JXF_CELLS"].split(","))
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
# no explicit init_process_group: the package creates the NCCL group itself from the torchrun environment
# (ParallelContext.from_environment) when the first manager needs the block runtime
s = H.make_setup(cells, bc=bc, gamma=1.4, length=1.0)
visc = os.environ.get("JXF_VISC", "0") == "1"
if visc:
    s.is_viscous_flux = s.is_heat_flux = True
    s.dynamic_viscosity, s.bulk_viscosity = 0.02, 0.003
    s.thermal_conductivity_model, s.prandtl_number, s.gas_constant = "PRANDTL", 0.71, 1.0
prims0 = H.smooth_ic(s, seed=21, amp=0.1)
case = {
  "general": {"case_name": "mg", "end_step": nsteps, "save_path": "./results"},
  "domain": {ax: {"cells": cells[i], "range": [0.0, 1.0]} for i, ax in enumerate("xyz")},
  "boundary_conditions": {f: {"type": s.bc[f]} for f in port.FACES},
  "initial_condition": {"rho": 1.0, "u": 0.0, "v": 0.0, "w": 0.0, "p": 1.0},
  "material_properties": {"equation_of_state": {"model": "IdealGas", "specific_heat_ratio": 1.4, "specific_gas_constant": 1.0}},
}
case["domain"]["decomposition"] = {"split_x": split[0], "split_y": split[1], "split_z": split[2]}
num = {"conservatives": {"halo_cells": 5, "time_integration": {"integrator": "RK3", "CFL": 0.5},
       "convective_fluxes": {"convective_solver": "GODUNOV", "godunov": {"riemann_solver": "HLLC", "signal_speed": "EINFELDT",
       "reconstruction_stencil": "WENO5-Z", "reconstruction_variable": "CHAR-PRIMITIVE"}}},
       "active_physics": {"is_convective_flux": True}, "output": {"logging": {"level": "NONE"}}}
if visc:
    num["active_physics"].update(is_viscous_flux=True, is_heat_flux=True)
    num["conservatives"]["dissipative_fluxes"] = {"reconstruction_stencil": "CENTRAL4", "derivative_stencil_center": "CENTRAL4",
                                                  "derivative_stencil_face": "CENTRAL4"}
    case["material_properties"]["transport"] = {"dynamic_viscosity": {"model": "CUSTOM", "value": 0.02}, "bulk_viscosity": 0.003,
                                                "thermal_conductivity": {"model": "PRANDTL", "prandtl_number": 0.71}}
im = InputManager(case, num)
init = InitializationManager(im)
assert dist.is_initialized()
rank = dist.get_rank()
active = [i for i in range(3) if cells[i] > 1]
user = prims0[[0] + [1 + i for i in active] + [4]]
buf = init.initialization(user_prime_init=user)
sim = SimulationManager(im)
sim.simulate(buf)
out = sim.final_buffers
di = im.domain_information
nh = 5
it = (slice(None),) + tuple(slice(nh, -nh) if cells[i] > 1 else slice(None) for i in range(3))
mine = out.simulation_buffers.material_fields.primitives[it].cpu().numpy()
# oracle on the global grid (single block), rank 0 gathers
gathered = [None] * dist.get_world_size()
dist.all_gather_object(gathered, (di.block_slices(rank), mine, out.time_control_variables.physical_timestep_size,
                                  out.time_control_variables.physical_simulation_time))
if rank == 0:
    glob = np.empty((5,) + cells)
    for sl, arr, _, _ in gathered:
        glob[(slice(None),) + sl] = arr
    p, c = port.initialize(user, s, from_user_buffer=True)
    dt = port.time_step_size(p, s); t = 0.0
    for _ in range(nsteps):
        t += dt
        p, c, dt = port.step(p, c, dt, s)
    ref = p[(slice(None),) + s.interior]
    err = H.rel_linf(glob, ref)
    dts = [g[2] for g in gathered]
    print("RESULT " + json.dumps({"err": err, "dt_err": max(abs(d - dt) / dt for d in dts), "t_err": abs(gathered[0][3] - t) / t}))
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.parametrize("visc", [0, 1])
@pytest.mark.parametrize("split,cells,bc", [((2, 1, 1), (32, 20, 36), "PERIODIC"), ((1, 2, 1), (20, 32, 36), "SYMMETRY"),
                                            ((1, 1, 2), (12, 16, 80), "PERIODIC"), ((2, 1, 1), (64, 24, 1), "ZEROGRADIENT"),
                                            ((1, 2, 1), (24, 40, 20), "ZEROGRADIENT")])
def test_two_blocks_match_single_block_oracle(split, cells, bc, visc, tmp_path):
    """visc=1: viscous + heat flux, i.e. the inter-block EDGE halos ride on the widened face slabs."""
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    worker = tmp_path / "worker.py"
    worker.write_text(WORKER)
    env = dict(os.environ, JXF_ROOT=ROOT, JXF_SPLIT=",".join(map(str, split)), JXF_BC=bc, JXF_STEPS="4",
               JXF_CELLS=",".join(map(str, cells)), JXF_VISC=str(visc))
    _run_worker(worker, env, 2)


def _run_worker(worker, env, nproc):
    port_no = 29500 + (os.getpid() % 200)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
                          "--master-addr", "127.0.0.1", "--master-port", str(port_no), str(worker)],
                         env=env, capture_output=True, text=True, timeout=600)
    lines = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")]
    assert lines, out.stdout[-2000:] + out.stderr[-4000:]
    res = json.loads(lines[-1][7:])
    assert res["err"] <= 1e-12 and res["dt_err"] <= 1e-12 and res["t_err"] <= 1e-12, res


@pytest.mark.parametrize("visc", [0, 1])
@pytest.mark.parametrize("split,cells,bc,nproc", [((2, 2, 1), (32, 28, 20), "SYMMETRY", 4), ((2, 2, 1), (24, 32, 16), "PERIODIC", 4),
                                                  ((1, 2, 2), (12, 24, 40), "ZEROGRADIENT", 4), ((2, 2, 2), (24, 20, 28), "SYMMETRY", 8)])
def test_pencil_and_block_decompositions(split, cells, bc, nproc, visc, tmp_path):
    """4 and 8 blocks: edges shared by two split axes (neighbour x neighbour), neighbour x physical edges, and
    with visc=1 the edge halos the dissipative stencils read there."""
    if _ngpus() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    worker = tmp_path / "worker.py"
    worker.write_text(WORKER)
    env = dict(os.environ, JXF_ROOT=ROOT, JXF_SPLIT=",".join(map(str, split)), JXF_BC=bc, JXF_STEPS="3",
               JXF_CELLS=",".join(map(str, cells)), JXF_VISC=str(visc))
    _run_worker(worker, env, nproc)


API_WORKER = r'''
import json, os, sys, signal
signal.alarm(300)
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["JXF_ROOT"])
from jaxfluids_b200 import InputManager, InitializationManager, SimulationManager
from tests import helpers as H

name = os.environ["JXF_FIXTURE"]
split = tuple(int(v) for v in os.environ["JXF_SPLIT"].split(","))
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
g, case, num = H.load_golden(name)
n = len(g["dt"])
case, num = json.loads(json.dumps(case)), json.loads(json.dumps(num))
case["general"]["end_step"] = n
case["general"]["end_time"] = 1e300
case["domain"]["decomposition"] = {"split_x": split[0], "split_y": split[1], "split_z": split[2]}
num.setdefault("output", {}).setdefault("logging", {})["level"] = "NONE"
im = InputManager(case, num)
buf = InitializationManager(im).initialization()
assert dist.is_initialized()
rank = dist.get_rank()
sim = SimulationManager(im)
assert sim.runtime.face_data, "this rank's block carries boundary data (jxf_set_face_data)"
sim.simulate(buf)
out = sim.final_buffers
s = H.setup_from_json(case, num)
di = im.domain_information
it = (slice(None),) + s.interior
mine = out.simulation_buffers.material_fields.primitives[it].cpu().numpy()
gathered = [None] * dist.get_world_size()
dist.all_gather_object(gathered, (di.block_slices(rank), mine, out.time_control_variables.physical_timestep_size))
if rank == 0:
    ref = g[f"prims_n{n}"][it]
    glob = np.empty_like(ref)
    for sl, arr, _ in gathered:
        glob[(slice(None),) + sl] = arr
    err = H.rel_linf(glob, ref)
    print("RESULT " + json.dumps({"err": err, "dt_err": max(abs(d - g["dt"][n - 1]) / g["dt"][n - 1] for _, _, d in gathered),
                                  "t_err": 0.0}))
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.parametrize("name,split", [("api/riemann2d_16x20_inflow_outflow_neumann_rk3", (2, 1, 1)),
                                        ("api/riemann2d_16x20_inflow_outflow_neumann_rk3", (1, 2, 1)),
                                        ("api/dmr_48x32_dirichlet_symmetry_south_rk3", (2, 1, 1)),
                                        ("api/heat2d_24x20_dirichlet_lambda_noconv_rk3", (2, 1, 1)),
                                        ("api/riemann2d_16x20_inflow_outflow_visc_rk3", (1, 2, 1)),
                                        ("api/cavity_24x20_wall_lambda_lid_visc_rk3", (2, 1, 1))])
def test_boundary_data_fixtures_on_two_blocks(name, split, tmp_path):
    """The reference's fixtures with boundary DATA (NEUMANN, SIMPLE_INFLOW / SIMPLE_OUTFLOW, space-dependent DIRICHLET and
    WALL velocities, a multi-type face) through the public API on TWO blocks: every block applies the data of its own
    outer faces in its kernels (per-block transverse cells), the shared face is exchanged; result = the reference's
    single-block fixture."""
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    worker = tmp_path / "worker.py"
    worker.write_text(API_WORKER)
    env = dict(os.environ, JXF_ROOT=ROOT, JXF_SPLIT=",".join(map(str, split)), JXF_FIXTURE=name)
    port_no = 29500 + (os.getpid() % 200)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port_no), str(worker)],
                         env=env, capture_output=True, text=True, timeout=600)
    lines = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")]
    assert lines, out.stdout[-2000:] + out.stderr[-4000:]
    res = json.loads(lines[-1][7:])
    assert res["err"] <= H.TOL_PRIMS_100 and res["dt_err"] <= 1e-10, res


@pytest.mark.parametrize("split,cells,bc,nproc", [((2, 1, 1), (32, 20, 36), "PERIODIC", 2), ((1, 2, 1), (20, 32, 36), "SYMMETRY", 2),
                                                  ((1, 1, 2), (12, 16, 80), "PERIODIC", 2), ((2, 1, 1), (64, 24, 1), "ZEROGRADIENT", 2),
                                                  ((2, 2, 1), (24, 32, 16), "PERIODIC", 4), ((2, 2, 2), (24, 20, 28), "SYMMETRY", 8)])
def test_peer_memory_halo_exchange(split, cells, bc, nproc, tmp_path):
    """JXF_PEER_HALO=1: the fused epilogue stores the halo images of shared faces straight into the neighbour's buffers
    (CUDA IPC over NVLink), flags instead of pack / NCCL / unpack -- same result as the single-block oracle."""
    if _ngpus() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    worker = tmp_path / "worker.py"
    worker.write_text(WORKER)
    env = dict(os.environ, JXF_ROOT=ROOT, JXF_SPLIT=",".join(map(str, split)), JXF_BC=bc, JXF_STEPS="4",
               JXF_CELLS=",".join(map(str, cells)), JXF_VISC="0", JXF_PEER_HALO="1")
    _run_worker(worker, env, nproc)
