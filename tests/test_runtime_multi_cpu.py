"""CPU, world_size 2, 4 and 8 over gloo: the WHOLE host side of the multi-block path -- InputManager -> InitializationManager ->
SimulationManager.simulate -> BlockRuntime with two blocks: NEIGHBOR faces, the 3-layer stage exchange with lazy halo
completion, the split first sweep of the overlap branch (interior range, then the strips next to the shared faces), the
MAX all-reduce of the step scalars ordered after the exchange -- with the CUDA solver replaced by an oracle-backed stand-in
on CPU tensors (the per-call contract of the kernels is what tests/test_gpu_parity.py and tests/test_gpu_multi.py check on
GPUs).  The two blocks' result must equal the single-block oracle on the global grid BIT FOR BIT: same arithmetic, only
the bookkeeping differs.  CUDA streams / events are replaced by no-ops (everything is synchronous on the CPU)."""
import json
import os
import subprocess
import sys

import pytest

from tests import helpers as H

ROOT = H.ROOT

WORKER = r'''
import copy, json, os, sys, signal
signal.alarm(300)
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["JXF_ROOT"])
from oracle import port
from tests import helpers as H
from tests.test_runtime_cpu import OracleSolver
import jaxfluids_b200.runtime as RT
from jaxfluids_b200 import InputManager, InitializationManager, SimulationManager


class _NoOp:                       # stand-in for torch.cuda.Stream / Event: the CPU run is synchronous
    def __init__(self, *a, **k): pass
    def wait_event(self, *a): pass
    def wait_stream(self, *a): pass
    def record(self, *a): pass
    def synchronize(self): pass
    def __enter__(self): return self
    def __exit__(self, *a): return False
torch.cuda.Stream = _NoOp
torch.cuda.Event = _NoOp
torch.cuda.stream = lambda s: _NoOp()
torch.cuda.current_stream = lambda *a, **k: _NoOp()
torch.cuda.synchronize = lambda *a, **k: None

FACE_AX = {0: 0, 1: 0, 2: 1, 3: 1, 4: 2, 5: 2}
CALLS = {"sweep_range": 0, "stage_tail": 0, "stage": 0, "pack": [], "unpack": []}


class OracleBlockSolver(OracleSolver):
    """One block of a decomposed grid: OracleSolver + the calls BlockRuntime makes only with neighbours."""

    def __init__(self, cfg):
        super().__init__(cfg)
        ref = OracleSolver.reference_setup
        s = self.setup
        s.cells = tuple(cfg.cells)
        s.active = tuple(i for i in range(3) if s.cells[i] > 1)
        s.inv_dx_override = tuple(ref.inv_dx)
        s.domain = tuple((0.0, float(ref.dx[i]) * s.cells[i]) for i in range(3))      # same dx (powers of two: exact)
        assert all(float(s.dx[i]) == float(ref.dx[i]) for i in s.active)
        self.active = s.active

    def _outer(self, base=None):       # NEIGHBOR faces are not the halo kernel's business
        s = copy.copy(base or self.setup)
        s.bc = {f: ("INACTIVE" if t == "NEIGHBOR" else t) for f, t in s.bc.items()}
        return s

    def _faces_only(self):
        return self._outer(super()._faces_only())

    def _physical_edges(self, p, c):
        """jxf_halo_fill_edges on a block: the reference's edge rule for edges between two PHYSICAL faces only; edges next
        to a shared face arrive with the widened slabs of the exchange (runtime._halo_update_with_edges)"""
        s = self.setup
        p2, c2 = port.edge_halo_fill(p, c, self._outer())
        nh = s.nh
        for edge in port.EDGES:
            fa, fb = edge.split("_")
            axa, axb = port.FACE_AXIS[fa], port.FACE_AXIS[fb]
            if axa not in s.active or axb not in s.active or "NEIGHBOR" in (s.bc[fa], s.bc[fb]):
                continue
            sl = [slice(None)] + list(s.interior)
            sl[1 + axa] = port._edge_range(fa, None, nh)
            sl[1 + axb] = port._edge_range(fb, None, nh)
            p[tuple(sl)] = p2[tuple(sl)]
            c[tuple(sl)] = c2[tuple(sl)]
        return p, c

    def halo_fill(self, prims, cons):
        s = self.setup
        with np.errstate(all="ignore"):
            p, c = port.halo_fill(prims.numpy(), cons.numpy(), self._faces_only())     # physical faces
            if s.is_dissipative and len(s.active) > 1:
                p, c = self._physical_edges(p, c)
        prims.copy_(torch.as_tensor(p))
        cons.copy_(torch.as_tensor(c))

    def halo_fill_edges(self, prims, cons):
        with np.errstate(all="ignore"):
            p, c = self._physical_edges(prims.numpy().copy(), cons.numpy().copy())
        prims.copy_(torch.as_tensor(p))
        cons.copy_(torch.as_tensor(c))

    def temperature(self, prims):
        with np.errstate(all="ignore"):
            return torch.as_tensor(port.temperature(prims.numpy(), self.setup))

    # -- face slabs ------------------------------------------------------------------------------------------------
    def _slab(self, face, layers, halo, ext=0):
        """ext (jxf_pack_face_ext): transverse widening over the nh halo cells -- bit 0 / 1: low / high side of the slower
        transverse axis, bit 2 / 3: of the faster one"""
        s = self.setup
        nh, ax = s.nh, FACE_AX[face]
        n = s.cells[ax]
        hi = face % 2 == 0
        if halo:
            rng = slice(nh + n, nh + n + layers) if hi else slice(nh - layers, nh)
        else:
            rng = slice(nh + n - layers, nh + n) if hi else slice(nh, nh + layers)
        sl = [slice(None)] + list(s.interior)
        sl[1 + ax] = rng
        t1, t2 = (1 if ax == 0 else 0), (1 if ax == 2 else 2)
        for bit, t in ((0, t1), (2, t2)):
            if s.cells[t] > 1:
                lo = nh - (nh if ext & (1 << bit) else 0)
                up = nh + s.cells[t] + (nh if ext & (2 << bit) else 0)
                sl[1 + t] = slice(lo, up)
        return tuple(sl)

    def face_slab_elems(self, face, ext=0, layers=None):
        s = self.setup
        layers = s.nh if layers is None else layers
        shape = np.empty(s.shape, dtype=np.int8)[self._slab(face, layers, False, ext)].shape
        return int(np.prod(shape))

    def pack_face(self, face, prims, buf, ext=0, layers=None):
        layers = self.setup.nh if layers is None else layers
        CALLS["pack"].append(layers)
        src = prims.numpy()[self._slab(face, layers, False, ext)]
        buf[:src.size].copy_(torch.as_tensor(np.ascontiguousarray(src).ravel()))

    def unpack_face(self, face, buf, prims, cons, ext=0, layers=None):
        layers = self.setup.nh if layers is None else layers
        CALLS["unpack"].append(layers)
        sl = self._slab(face, layers, True, ext)
        shape = prims.numpy()[sl].shape
        got = buf.numpy()[:int(np.prod(shape))].reshape(shape)
        prims.numpy()[sl] = got
        with np.errstate(all="ignore"):
            cons.numpy()[sl] = port.cons_from_prims(got, self.setup.gamma)

    # -- sweeps ----------------------------------------------------------------------------------------------------
    def sweep_range(self, axis, lo, hi, prims, rhs, accumulate=False):
        CALLS["sweep_range"] += 1
        with np.errstate(all="ignore"):
            r = port.rhs_axis(prims.numpy(), axis, self.setup)
        sl = [slice(None)] * 4
        sl[1 + axis] = slice(lo, hi)
        sl = tuple(sl)
        out = rhs.numpy()
        out[sl] = (out[sl] + r[sl]) if accumulate else (0.0 + r[sl])

    def stage_tail(self, k, first_done, p_in, p_out, c_in, c_n, c_out, rhs, dt, red, reduce=False, fill_halo=True):
        CALLS["stage_tail" if first_done else "stage"] += 1
        assert fill_halo
        s, rk = self.setup, port.RK[self.setup.integrator]
        with np.errstate(all="ignore"):
            r = rhs.numpy().copy() if first_done else 0.0
            for axis in s.active[first_done:]:
                r = r + port.rhs_axis(p_in.numpy(), axis, s)
            cons = c_in.numpy()
            if k > 0:
                a, b = rk["blend"][k - 1]
                cons = a * cons + b * c_n.numpy()
            cons = cons.copy()
            sl = (slice(None),) + s.interior
            cons[sl] = cons[sl] + (float(dt.item()) * rk["dt_mult"][k]) * r
            prims = port.prims_from_cons(cons, s.gamma)
        c_out.copy_(torch.as_tensor(cons))
        p_out.copy_(torch.as_tensor(prims))
        self.halo_fill(p_out, c_out)
        if reduce:
            self.reduce(p_out, red)

    def stage(self, k, p_in, p_out, c_in, c_n, c_out, rhs, dt, red, reduce=False, fill_halo=True):
        self.stage_tail(k, 0, p_in, p_out, c_in, c_n, c_out, rhs, dt, red, reduce=reduce, fill_halo=fill_halo)

    # -- step scalars: the block's {max sum(|u| + c), min rho, min p}; the runtime MAX-all-reduces them ----------------
    def reduce_reset(self, red):
        red.copy_(torch.tensor([0.0, float("inf"), float("inf")], dtype=torch.float64))

    def reduce(self, prims, red):
        s = self.setup
        pi = prims.numpy()[(slice(None),) + s.interior]
        c = port.speed_of_sound(pi[4], pi[0], s.gamma)
        acc = 0.0
        for i in s.active:
            acc = acc + (np.abs(pi[1 + i]) + c)
        red.copy_(torch.tensor([max(float(red[0]), float(np.max(acc))), min(float(red[1]), float(np.min(pi[0]))),
                                min(float(red[2]), float(np.min(pi[4])))], dtype=torch.float64))

    def finish_step(self, red, dt, time, info):
        s = self.setup
        if time is not None:
            time += dt
        if info is not None:
            info.copy_(red)
        d = s.dx_min / (np.float64(red[0].item()) + port.EPS)          # time_step_size.py:15-157, global maximum
        if s.is_dissipative:           # :111-135 with constant mu / lambda: max(mu / rho) = mu / min(rho) (finish_step_kernel)
            one = np.ones(1)
            dx2 = s.dx_min * s.dx_min
            rho_min = np.float64(red[1].item())
            if s.is_viscous_flux:
                d = np.minimum(d, 3.0 / 14.0 * dx2 / (port._dynamic_viscosity(one, s)[0] / rho_min + port.EPS))
            if s.is_heat_flux:
                d = np.minimum(d, 0.1 * dx2 / (port._thermal_conductivity(one, s)[0] / (rho_min * s.cp) + port.EPS))
        dt.fill_(float(d * s.cfl))
        self.reduce_reset(red)


split = tuple(int(v) for v in os.environ["JXF_SPLIT"].split(","))
bc = os.environ["JXF_BC"]
nsteps = int(os.environ["JXF_STEPS"])
cells = tuple(int(v) for v in os.environ["JXF_CELLS"].split(","))
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
OracleSolver.reference_setup = s
RT.BlockSolver = OracleBlockSolver
im = InputManager(case, num)
init = InitializationManager(im)
assert dist.is_initialized() and dist.get_backend() == "gloo"
rank = dist.get_rank()
active = [i for i in range(3) if cells[i] > 1]
user = prims0[[0] + [1 + i for i in active] + [4]]
buf = init.initialization(user_prime_init=user)
sim = SimulationManager(im)
rt = sim.runtime
assert rt.neighbors and rt.stage_layers == (5 if visc else int(os.environ.get("JXF_EXCHANGE_LAYERS", "3").replace("full", "5")))
sim.simulate(buf)
out = sim.final_buffers
di = im.domain_information
nh = 5
it = (slice(None),) + tuple(slice(nh, -nh) if cells[i] > 1 else slice(None) for i in range(3))
full = out.simulation_buffers.material_fields.primitives.numpy()
mine = full[it]
# every halo layer of the returned buffers is up to date (complete_halos): the halo next to a shared face equals the
# neighbour's interior; checked through the global array below
gathered = [None] * dist.get_world_size()
dist.all_gather_object(gathered, (di.block_slices(rank), mine, out.time_control_variables.physical_timestep_size,
                                  out.time_control_variables.physical_simulation_time, full, dict(CALLS),
                                  sorted(rt.neighbors), bool(rt.overlap)))
if rank == 0:
    glob = np.empty((5,) + cells)
    for g in gathered:
        glob[(slice(None),) + g[0]] = g[1]
    p, c = port.initialize(user, s, from_user_buffer=True)
    dt = port.time_step_size(p, s); t = 0.0
    for _ in range(nsteps):
        t += dt
        p, c, dt = port.step(p, c, dt, s)
    ref = p[(slice(None),) + s.interior]
    # halos of the returned block buffers against the oracle's halo'd global buffer (all nh layers, shared faces included)
    halo_ok = True
    for g in gathered:
        sl, blk = g[0], g[4]
        idx = [np.arange(sl[i].start, sl[i].stop + 2 * nh) if cells[i] > 1 else np.arange(1) for i in range(3)]
        want = p[np.ix_(np.arange(5), *idx)]
        m = np.zeros(blk.shape[1:], bool)                      # face halos only (corners / edges are not defined here)
        for ax in active:
            e = [slice(nh, -nh) if (cells[i] > 1) else slice(None) for i in range(3)]
            for side in (slice(0, nh), slice(-nh, None)):
                e2 = list(e); e2[ax] = side; m[tuple(e2)] = True
        halo_ok = halo_ok and np.array_equal(blk[:, m], want[:, m])
    print("RESULT " + json.dumps({"equal": bool(np.array_equal(glob, ref)), "halo_equal": bool(halo_ok),
                                  "dt_equal": all(g[2] == dt for g in gathered), "t_equal": gathered[0][3] == t,
                                  "calls": [g[5] for g in gathered], "neighbors": [g[6] for g in gathered],
                                  "overlap": [g[7] for g in gathered]}))
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.parametrize("split,cells,bc,layers,visc", [
    ((2, 1, 1), (32, 16, 8), "PERIODIC", "3", 0), ((1, 2, 1), (8, 32, 16), "SYMMETRY", "3", 0),
    ((1, 1, 2), (8, 16, 32), "PERIODIC", "3", 0), ((2, 1, 1), (32, 16, 1), "ZEROGRADIENT", "full", 0),
    # viscous + heat flux: the exchange also carries the EDGE halos next to the shared faces (widened slabs, axis by axis)
    ((2, 1, 1), (32, 16, 8), "PERIODIC", "3", 1), ((1, 2, 1), (16, 32, 1), "SYMMETRY", "3", 1),
    ((1, 1, 2), (8, 16, 32), "ZEROGRADIENT", "3", 1),
    # four blocks (pencils): two split axes -- with the dissipative fluxes the edge halos at the line where four blocks meet
    # come from the DIAGONAL neighbour in two hops (x exchange, then y exchange over the halos x just filled)
    ((2, 2, 1), (16, 16, 8), "PERIODIC", "3", 0), ((2, 2, 1), (16, 16, 8), "PERIODIC", "3", 1),
    ((2, 2, 1), (32, 16, 1), "SYMMETRY", "3", 1),
    # eight blocks: every block has three neighbours, the edge halos of the dissipative stencils cross all three axes
    ((2, 2, 2), (16, 16, 16), "PERIODIC", "3", 0), ((2, 2, 2), (16, 16, 16), "SYMMETRY", "3", 1)])
def test_two_blocks_through_the_host_runtime_equal_the_single_block_oracle(split, cells, bc, layers, visc, tmp_path):
    worker = tmp_path / "worker.py"
    worker.write_text(WORKER)
    env = dict(os.environ, JXF_ROOT=ROOT, JXF_SPLIT=",".join(map(str, split)), JXF_BC=bc, JXF_STEPS="2", JXF_VISC=str(visc),
               JXF_CELLS=",".join(map(str, cells)), JXF_EXCHANGE_LAYERS=layers, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="2")
    port_no = 29700 + (os.getpid() + sum(cells) + len(bc)) % 200
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={split[0] * split[1] * split[2]}",
                          "--master-addr", "127.0.0.1", "--master-port", str(port_no), str(worker)],
                         env=env, capture_output=True, text=True, timeout=900)
    lines = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")]
    assert lines, out.stdout[-2000:] + out.stderr[-5000:]
    res = json.loads(lines[-1][7:])
    assert res["equal"] and res["dt_equal"] and res["t_equal"], res
    assert res["halo_equal"], "halos of the returned buffers are not complete"
    nl = 5 if (layers == "full" or visc) else int(layers)
    for calls, nbrs, overlap in zip(res["calls"], res["neighbors"], res["overlap"]):
        assert len(nbrs) >= 1
        if visc:                         # full slabs every stage, no overlap on this path
            assert not overlap and set(calls["pack"]) == {5} and calls["stage_tail"] == 0
            continue
        assert overlap
        # between stages only `nl` layers travel; the hand-over to the user ships all five
        assert set(calls["pack"]) <= {nl, 5} and nl in calls["pack"] and 5 in calls["pack"]
        assert calls["pack"] == calls["unpack"] or sorted(calls["pack"]) == sorted(calls["unpack"])
        # the overlap branch split the first sweep whenever an exchange was in flight
        assert calls["stage_tail"] > 0 and calls["sweep_range"] >= calls["stage_tail"]
