"""CPU: the WHOLE host side of the boundary-data path -- InputManager -> InitializationManager ->
SimulationManager.simulate -> BlockRuntime (stage by stage, halo kernel, host boundary data, edge fill, buffer rotation,
time step) -- with the CUDA solver replaced by an oracle-backed stand-in on CPU tensors.  What the stand-in does per call
is what the kernels are tested to do on the GPU (tests/test_gpu_parity.py); this test checks that the runtime calls them
in an order and with buffers that reproduce the reference's fixtures (tests/golden/api/), bit for bit."""
import copy
import json

import numpy as np
import pytest
import torch

from oracle import port
from tests import helpers as H


class OracleSolver:
    """The BlockSolver calls BlockRuntime makes on this path, computed by oracle/port.py on CPU tensors.  Boundary types
    are the ones the KERNELS are configured with (cfg.bc / cfg.dirichlet / cfg.wall_velocity: placeholders included)."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.device = torch.device("cpu")
        self.setup = copy.copy(OracleSolver.reference_setup)
        s = self.setup
        s.bc = dict(cfg.bc)
        s.bc_multi = {}
        s.dirichlet = {f: tuple(v) for f, v in cfg.dirichlet.items()}
        s.wall_velocity = {f: tuple(v) for f, v in cfg.wall_velocity.items()}
        s.bc_values = {}
        self.active = s.active
        self.stages = port.RK[s.integrator]["stages"]
        self._last = None

    # allocation
    def new_field(self, fill=None):
        t = torch.empty(self.cfg.shape, dtype=torch.float64)
        return t.fill_(fill) if fill is not None else t

    def new_rhs(self):
        return torch.empty(self.cfg.rhs_shape, dtype=torch.float64)

    def new_scalars(self, n=1, value=0.0):
        return torch.full((n,), value, dtype=torch.float64)

    def new_red(self):
        return torch.tensor([0.0, float("inf"), float("inf")], dtype=torch.float64)

    def bind_timestep(self, dt):
        pass

    def set_face_data(self, face, ops, data, mask=None):
        """jxf_set_face_data: kept per face, applied after the base rule by halo_fill (helpers.apply_face_data_numpy)"""
        self.face_data = getattr(self, "face_data", {})
        self.face_data[port.FACES[face]] = (ops, data, mask)

    # "kernels"
    def cons_from_prims(self, prims, cons):
        with np.errstate(all="ignore"):
            cons.copy_(torch.as_tensor(port.cons_from_prims(prims.numpy(), self.setup.gamma)))

    def _faces_only(self):
        s = copy.copy(self.setup)
        s.is_viscous_flux = s.is_heat_flux = False
        return s

    def halo_fill(self, prims, cons):
        s = self.setup
        with np.errstate(all="ignore"):
            p, c = port.halo_fill(prims.numpy(), cons.numpy(), self._faces_only())     # the faces' base rules
            p, c = H.apply_face_data_numpy(getattr(self, "face_data", {}), p, c, s.cells, s.nh, s.gamma)
            if s.is_dissipative and len(s.active) > 1:                               # then the edges
                p, c = port.edge_halo_fill(p, c, s)
        prims.copy_(torch.as_tensor(p))
        cons.copy_(torch.as_tensor(c))

    def halo_fill_edges(self, prims, cons):
        with np.errstate(all="ignore"):
            p, c = port.edge_halo_fill(prims.numpy().copy(), cons.numpy().copy(), self.setup)
        prims.copy_(torch.as_tensor(p))
        cons.copy_(torch.as_tensor(c))

    def temperature(self, prims):
        return None

    def reduce_reset(self, red):
        self._last = None

    def reduce(self, prims, red):
        self._last = prims.numpy().copy()

    def stage(self, k, p_in, p_out, c_in, c_n, c_out, rhs, dt, red, reduce=False, fill_halo=True):
        assert fill_halo, "the stage carries the halo images (fused in the kernels; here: halo_fill after the update)"
        s, rk = self.setup, port.RK[self.setup.integrator]
        with np.errstate(all="ignore"):
            r = port.compute_rhs(p_in.numpy(), s, c_in.numpy(), float(dt.item()))
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
            self._last = p_out.numpy().copy()

    def step_fused(self, pa, pb, ca, cb, rhs, dt, time, red, info, fill_halo=True):
        """jxf_step_fused: all stages on the ping-pong buffers, then the step scalars; returns the buffer parity"""
        pr, cur = [pa, pb], 0
        for k in range(self.stages):
            last = k == self.stages - 1
            self.stage(k, pr[cur], pr[cur ^ 1], ca if k == 0 else cb, ca, ca if last else cb, rhs, dt, red, reduce=last,
                       fill_halo=fill_halo)
            cur ^= 1
        self.finish_step(red, dt, time, info)
        return cur

    def finish_step(self, red, dt, time, info):
        s = self.setup
        if time is not None:
            time += dt
        mr, mp = port.positivity_info(self._last, s)
        dt.fill_(port.time_step_size(self._last, s))
        info.copy_(torch.tensor([0.0, mr, mp], dtype=torch.float64))


@pytest.mark.parametrize("name", H.api_golden_names())
def test_public_api_on_cpu_with_oracle_backed_solver(name, monkeypatch):
    import jaxfluids_b200.runtime as RT
    from jaxfluids_b200 import InputManager, InitializationManager, SimulationManager
    g, case, num = H.load_golden(name)
    n = len(g["dt"])
    case, num = json.loads(json.dumps(case)), json.loads(json.dumps(num))
    case["general"]["end_step"] = n
    case["general"]["end_time"] = 1e300
    num.setdefault("output", {}).setdefault("logging", {})["level"] = "NONE"
    s = H.setup_from_json(case, num)
    OracleSolver.reference_setup = s
    monkeypatch.setattr(RT, "BlockSolver", OracleSolver)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    im = InputManager(case, num)
    buffers = InitializationManager(im).initialization()
    m = H.defined_mask(s)
    mf = buffers.simulation_buffers.material_fields
    assert np.array_equal(mf.primitives.numpy()[:, m], g["prims0_halo"][:, m])
    assert np.array_equal(mf.conservatives.numpy()[:, m], g["cons0_halo"][:, m])
    assert buffers.time_control_variables.physical_timestep_size == float(g["dt0"])
    sim = SimulationManager(im)
    assert sim.runtime.face_data and not sim.runtime._host_halo
    sim.simulate(buffers)
    out = sim.final_buffers
    tcv = out.time_control_variables
    assert tcv.simulation_step == n
    assert tcv.physical_timestep_size == g["dt"][n - 1]
    assert tcv.physical_simulation_time == g["time"][n - 1]
    omf = out.simulation_buffers.material_fields
    assert np.array_equal(omf.primitives.numpy()[:, m], g[f"prims_n{n}"][:, m], equal_nan=True)
    assert np.array_equal(omf.conservatives.numpy()[:, m], g[f"cons_n{n}"][:, m], equal_nan=True)
    pos = out.step_information.positivity[-1]
    assert (pos.min_density, pos.min_pressure) == (g["min_density"][n - 1], g["min_pressure"][n - 1])
