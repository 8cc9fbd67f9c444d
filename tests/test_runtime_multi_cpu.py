"""CPU, world_size 2 over gloo: the WHOLE host side of the multi-block path -- InputManager -> InitializationManager ->
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

    def halo_fill(self, prims, cons):
        keep = self.setup
        self.setup = self._outer()
        try:
            super().halo_fill(prims, cons)
        finally:
            self.setup = keep

    # -- face slabs ------------------------------------------------------------------------------------------------
    def _slab(self, face, layers, halo):
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
        return tuple(sl)

    def face_slab_elems(self, face, ext=0, layers=None):
        assert ext == 0
        s = self.setup
        layers = s.nh if layers is None else layers
        t = [s.cells[i] for i in range(3) if i != FACE_AX[face]]
        return 5 * layers * t[0] * t[1]

    def pack_face(self, face, prims, buf, ext=0, layers=None):
        layers = self.setup.nh if layers is None else layers
        CALLS["pack"].append(layers)
        src = prims.numpy()[self._slab(face, layers, halo=False)]
        buf[:src.size].copy_(torch.as_tensor(np.ascontiguousarray(src).ravel()))

    def unpack_face(self, face, buf, prims, cons, ext=0, layers=None):
        layers = self.setup.nh if layers is None else layers
        CALLS["unpack"].append(layers)
        sl = self._slab(face, layers, halo=True)
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
        dt.fill_(float(d * s.cfl))
        self.reduce_reset(red)


split = tuple(int(v) for v in os.environ["JXF_SPLIT"].split(","))
bc = os.environ["JXF_BC"]
nsteps = int(os.environ["JXF_STEPS"])
cells = tuple(int(v) for v in os.environ["JXF_CELLS"].split(","))
s = H.make_setup(cells, bc=bc, gamma=1.4, length=1.0)
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
assert rt.neighbors and rt.stage_layers == int(os.environ.get("JXF_EXCHANGE_LAYERS", "3").replace("full", "5"))
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


@pytest.mark.parametrize("split,cells,bc,layers", [((2, 1, 1), (32, 16, 8), "PERIODIC", "3"), ((1, 2, 1), (8, 32, 16), "SYMMETRY", "3"),
                                                   ((1, 1, 2), (8, 16, 32), "PERIODIC", "3"), ((2, 1, 1), (32, 16, 1), "ZEROGRADIENT", "full")])
def test_two_blocks_through_the_host_runtime_equal_the_single_block_oracle(split, cells, bc, layers, tmp_path):
    worker = tmp_path / "worker.py"
    worker.write_text(WORKER)
    env = dict(os.environ, JXF_ROOT=ROOT, JXF_SPLIT=",".join(map(str, split)), JXF_BC=bc, JXF_STEPS="2",
               JXF_CELLS=",".join(map(str, cells)), JXF_EXCHANGE_LAYERS=layers, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="2")
    port_no = 29700 + (os.getpid() + sum(cells) + len(bc)) % 200
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port_no), str(worker)],
                         env=env, capture_output=True, text=True, timeout=900)
    lines = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")]
    assert lines, out.stdout[-2000:] + out.stderr[-5000:]
    res = json.loads(lines[-1][7:])
    assert res["equal"] and res["dt_equal"] and res["t_equal"], res
    assert res["halo_equal"], "halos of the returned buffers are not complete"
    nl = 5 if layers == "full" else int(layers)
    for calls, nbrs, overlap in zip(res["calls"], res["neighbors"], res["overlap"]):
        assert overlap and len(nbrs) >= 1
        # between stages only `nl` layers travel; the hand-over to the user ships all five
        assert set(calls["pack"]) <= {nl, 5} and nl in calls["pack"] and 5 in calls["pack"]
        assert calls["pack"] == calls["unpack"] or sorted(calls["pack"]) == sorted(calls["unpack"])
        # the overlap branch split the first sweep whenever an exchange was in flight
        assert calls["stage_tail"] > 0 and calls["sweep_range"] >= calls["stage_tail"]
