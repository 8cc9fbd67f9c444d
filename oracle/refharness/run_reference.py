"""Run the UNMODIFIED reference sources (tumaer/JAXFLUIDS under /root/reference/src)
on the NumPy-backed `jax` stand-in and record what the hot path produces.

TEST INFRASTRUCTURE ONLY.  Works only in the build container (needs
/root/reference); the GPU box never runs this.  `make_goldens.py` uses it to
write the fixtures under tests/golden/, and tests/test_oracle_pinning.py uses it
(when /root/reference exists) to pin oracle/port.py against the reference.

What is recorded per RK stage (hooks are wrappers around bound methods, the
reference code itself is untouched):
  * rhs       <- SpaceSolver.compute_rhs            (solvers/space_solver.py:151)
  * cons/prims after HaloManager.perform_halo_update_material
                                                    (halos/halo_manager.py:146)
and per step: dt (time_step_size.py:15), t, min rho / min p
(positivity_handler.py:245-254) and the interior sums of the conservatives.
"""
from __future__ import annotations

import copy
import json
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_SRC = os.environ.get("JXF_REFERENCE_SRC", "/root/reference/src")
REFERENCE_EXAMPLES = os.path.join(os.path.dirname(REFERENCE_SRC), "examples")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "jaxfluids"))


def _activate():
    """Put stand-in + stubs + reference on sys.path (idempotent)."""
    sys.dont_write_bytecode = True
    for p in (REFERENCE_SRC, os.path.join(HERE, "stubs"), os.path.join(HERE, "standin")):
        if p not in sys.path:
            sys.path.insert(0, p)
    warnings.filterwarnings("ignore", category=SyntaxWarning)
    import jax  # noqa: F401  (the stand-in)
    assert "standin" in jax.__file__, "a real jax shadowed the stand-in"


CASES = {
    "sod": ("examples_1D/02_sod_shock_tube", "sod.json"),
    "riemann2d": ("examples_2D/07_riemann_problem", "riemann2D.json"),
    "tgv": ("examples_3D/01_tgv", "tgv.json"),
    "cavity": ("examples_2D/03_lid_driven_cavity", "lid_driven_cavity.json"),   # WALL x4, WENO5-JS, viscous, limiter, nh 4
    "rti": ("examples_2D/04_rayleigh_taylor_instability", "rti.json"),           # DIRICHLET N/S, gravity, limiter
    "heat1d": ("examples_1D/08_heat_equation", "heat_equation.json"),            # heat flux only (no convective flux)
    "rarefaction": ("examples_1D/04_double_rarefaction", "double_rarefaction.json"),  # flux limiter SIMPLE + interp. limiter
    "heat2d": ("examples_2D/06_heat_equation", "heat_equation.json"),                  # heat flux only, DIRICHLET x4, p(x) at north
    "dmr": ("examples_2D/08_double_mach_reflection", "double_mach_reflection.json"),   # south: DIRICHLET | SYMMETRY by x, limiter
    "lax": ("examples_1D/03_lax_shock_tube", "lax.json"),                             # FLUX-SPLITTING ROE + WENO6-CU
    "woodward": ("examples_1D/06_woodward_shock_tube", "woodward_shock_tube.json"),   # FLUX-SPLITTING ROE + WENO5-Z, SYMMETRY
}


def load_case(name: str):
    """Return (case_dict, numerical_dict) of a shipped example, unmodified."""
    d, f = CASES[name]
    with open(os.path.join(REFERENCE_EXAMPLES, d, f)) as fh:
        case = json.load(fh)
    with open(os.path.join(REFERENCE_EXAMPLES, d, "numerical_setup.json")) as fh:
        num = json.load(fh)
    return case, num


def customize(case, num, cells=None, bc=None, recon=None, riemann=None, integrator=None, stencil=None,
              signal_speed=None, positivity=None, initial_condition=None, flux_splitting=None, frozen_state=None,
              boundary_conditions=None):
    case, num = copy.deepcopy(case), copy.deepcopy(num)
    if cells is not None:
        for ax, n in zip("xyz", cells):
            if n is not None:
                case["domain"][ax]["cells"] = int(n)
    if bc is not None:
        for face in ("east", "west", "north", "south", "top", "bottom"):
            if isinstance(case["boundary_conditions"][face], dict) and \
                    case["boundary_conditions"][face]["type"] not in ("INACTIVE", "DIRICHLET", "WALL"):
                case["boundary_conditions"][face] = {"type": bc}
    if boundary_conditions is not None:       # face -> the face's whole entry (type + callables)
        for face, entry in boundary_conditions.items():
            case["boundary_conditions"][face] = copy.deepcopy(entry)
    cf = num["conservatives"]["convective_fluxes"]
    if flux_splitting is not None:           # convective_solver FLUX-SPLITTING with this eigenvalue choice
        cf["convective_solver"] = "FLUX-SPLITTING"
        cf.setdefault("flux_splitting", {})
        cf["flux_splitting"]["flux_splitting"] = flux_splitting
        cf["flux_splitting"].setdefault("reconstruction_stencil",
                                        cf.get("godunov", {}).get("reconstruction_stencil", "WENO5-Z"))
        if stencil is not None:
            cf["flux_splitting"]["reconstruction_stencil"] = stencil
    g = cf.setdefault("godunov", {})
    if recon is not None:
        g["reconstruction_variable"] = recon
    if riemann is not None:
        g["riemann_solver"] = riemann
    if stencil is not None:
        g["reconstruction_stencil"] = stencil
    if signal_speed is not None:
        g["signal_speed"] = signal_speed
    if frozen_state is not None:             # of the block the selected solver reads
        (cf["flux_splitting"] if cf.get("convective_solver") == "FLUX-SPLITTING" else g)["frozen_state"] = frozen_state
    if integrator is not None:
        num["conservatives"]["time_integration"]["integrator"] = integrator
    if positivity is not None:
        num["conservatives"]["positivity"] = dict(positivity)
    if initial_condition is not None:
        case["initial_condition"].update(initial_condition)
    # keep the reference from writing anything / printing the banner
    num.setdefault("output", {})
    num["output"].setdefault("logging", {})
    num["output"]["logging"]["level"] = "NONE"
    return case, num


class ReferenceRun:
    """Drives InputManager -> InitializationManager -> SimulationManager._do_integration_step."""

    def __init__(self, case: dict, num: dict, user_prime_init=None):
        _activate()
        import numpy as np
        from jaxfluids import InputManager, InitializationManager, SimulationManager

        self.np = np
        self.input_manager = InputManager(case, num)
        self.init_manager = InitializationManager(self.input_manager)
        self.sim = SimulationManager(self.input_manager)
        if user_prime_init is not None:
            self.buffers = self.init_manager.initialization(user_prime_init=np.asarray(user_prime_init))
        else:
            self.buffers = self.init_manager.initialization()
        self.records = []  # one dict per step
        self._cur = None
        self._install_hooks()

    # -- hooks -----------------------------------------------------------
    def _install_hooks(self):
        np = self.np
        ss, hm = self.sim.space_solver, self.sim.halo_manager
        orig_rhs, orig_halo = ss.compute_rhs, hm.perform_halo_update_material

        def rhs_hook(*a, **k):
            out = orig_rhs(*a, **k)
            if self._cur is not None:
                self._cur["rhs"].append(np.array(out[0].euler_buffers.conservatives))
            return out

        def halo_hook(*a, **k):
            prims, cons = orig_halo(*a, **k)
            if self._cur is not None:
                self._cur["prims"].append(np.array(prims))
                self._cur["cons"].append(np.array(cons))
            return prims, cons

        ss.compute_rhs = rhs_hook
        hm.perform_halo_update_material = halo_hook

    # -- accessors -------------------------------------------------------
    @property
    def material_fields(self):
        return self.buffers.simulation_buffers.material_fields

    @property
    def primitives(self):
        return self.np.array(self.material_fields.primitives)

    @property
    def conservatives(self):
        return self.np.array(self.material_fields.conservatives)

    @property
    def dt(self):
        return float(self.buffers.time_control_variables.physical_timestep_size)

    @property
    def time(self):
        return float(self.buffers.time_control_variables.physical_simulation_time)

    def interior(self, a):
        di = self.sim.domain_information
        nhx, nhy, nhz = di.domain_slices_conservatives
        return a[..., nhx, nhy, nhz]

    def default_ml_setup(self):
        """The (empty) MachineLearningSetup the reference builds from simulate()'s defaults
        (simulation_manager.py:189-190, :582)."""
        from jaxfluids.data_types.ml_buffers import CallablesSetup, ParametersSetup, combine_callables_and_params
        return combine_callables_and_params(CallablesSetup(), ParametersSetup())

    def compute_rhs(self, prims=None, cons=None, dt=None):
        """One SpaceSolver.compute_rhs evaluation on the current (or given) state (dt: the physical time step
        size handed to the positivity flux limiter; default: the current one)."""
        mf = self.material_fields
        prims = mf.primitives if prims is None else prims
        cons = mf.conservatives if cons is None else cons
        save, self._cur = self._cur, None
        out = self.sim.space_solver.compute_rhs(cons, prims, mf.temperature, 0.0, self.dt if dt is None else dt,
                                                ml_setup=self.default_ml_setup())
        self._cur = save
        return self.np.array(out[0].euler_buffers.conservatives)

    def step(self, record_stages=True):
        np = self.np
        tcv = self.buffers.time_control_variables
        rec = {"dt_used": float(tcv.physical_timestep_size), "rhs": [], "prims": [], "cons": []}
        self._cur = rec
        cfp = self.sim.compute_control_flow_params(tcv, self.buffers.step_information)
        # the defaults SimulationManager.simulate passes (simulation_manager.py:189-190)
        from jaxfluids.data_types.ml_buffers import CallablesSetup, ParametersSetup
        self.buffers, _ = self.sim._do_integration_step(self.buffers, cfp, ParametersSetup(), CallablesSetup())
        self._cur = None
        tcv = self.buffers.time_control_variables
        rec["dt_next"] = float(tcv.physical_timestep_size)
        rec["time"] = float(tcv.physical_simulation_time)
        pos = self.buffers.step_information.positivity
        if pos:
            rec["min_density"] = float(pos[-1].min_density)
            rec["min_pressure"] = float(pos[-1].min_pressure)
        cons_int = self.interior(self.conservatives)
        rec["totals"] = np.array([cons_int[v].sum() for v in range(5)])
        if not record_stages:
            rec["rhs"], rec["prims"], rec["cons"] = [], [], []
        self.records.append(rec)
        return rec


if __name__ == "__main__":
    import numpy as np
    case, num = customize(*load_case("sod"), cells=(1000, None, None))
    run = ReferenceRun(case, num)
    print("dt0", run.dt, "expected", 0.5 * 1e-3 / np.sqrt(1.4))
    for _ in range(3):
        r = run.step()
        print(r["time"], r["dt_next"], r["totals"][0] / 1000, r.get("min_density"))
