"""InputManager: reads a JAX-Fluids case setup and numerical setup (paths to .json
files or dicts) and validates the keys the convective hot path uses.

Mirrors input/input_manager.py:32-116 of the reference for the path's subset:
the same JSON schema, the same defaults, the same option strings
(registries.py).  Options outside the path raise NotImplementedError naming the
JSON path; malformed values raise the reference's consistency AssertionError.
"""
from __future__ import annotations

import copy
import json
import os
from typing import Any, Callable, Dict, NamedTuple, Optional, Tuple, Union

import numpy as np

from . import registries as R
from .domain_information import DomainInformation

FACES = ("east", "west", "north", "south", "top", "bottom")          # domain/__init__.py:5-7
FACE_AXIS = {"east": 0, "west": 0, "north": 1, "south": 1, "top": 2, "bottom": 2}
AXES = ("x", "y", "z")


def _assert(cond, msg, setup):
    assert cond, f"Consistency error in {setup} setup file. {msg}"


def read_json_setup(setup: Union[str, Dict], name: str) -> Dict:
    """input/input_manager.py:341-366."""
    if isinstance(setup, dict):
        return copy.deepcopy(setup)
    if isinstance(setup, (str, os.PathLike)):
        path = os.fspath(setup)
        assert os.path.isfile(path), f"Consistency error reading {name} file. {path} does not exist."
        with open(path) as fh:
            return json.load(fh)
    raise AssertionError(f"Consistency error reading {name} file. {name} must be a path or a dictionary.")


def get_setup_value(d: Dict, key: str, path: str, types, is_optional: bool, default_value: Any = None,
                    possible_string_values=None, numerical_value_condition=None, setup="numerical"):
    """input/setup_reader.py (_get_setup_value): same checks, same messages."""
    def check(v):
        _assert(isinstance(v, types), f"Key {path} must be of types {types}, but is of type {type(v)}.", setup)
        if possible_string_values is not None and isinstance(v, str):
            _assert(v in possible_string_values,
                    f"Value of {path:s} must be in {possible_string_values} if value is of type str.", setup)
        if numerical_value_condition is not None and isinstance(v, (int, float)):
            op, ref = numerical_value_condition
            ok = {">": v > ref, "<": v < ref, ">=": v >= ref, "<=": v <= ref}[op]
            _assert(ok, f"Value of {path} must be {op:s} {str(ref):s}.", setup)
    if key in d:
        v = d[key]
        if not is_optional:
            check(v)
        elif v is not None:
            check(v)
        else:
            v = default_value
        return v
    if is_optional:
        return default_value
    _assert(False, f"Key {key:s} is not optional, but missing {path:s}.", setup)


# ---------------------------------------------------------------------------
# setup containers (names follow data_types/numerical_setup, data_types/case_setup)
# ---------------------------------------------------------------------------
class TimeIntegrationSetup(NamedTuple):
    integrator: str
    CFL: float
    fixed_timestep: Any


class HighOrderGodunovSetup(NamedTuple):
    riemann_solver: str
    signal_speed: str
    reconstruction_stencil: str
    reconstruction_variable: str
    frozen_state: str


class FluxSplittingSetup(NamedTuple):
    """data_types/numerical_setup/conservatives.py:44-48."""
    flux_splitting: str
    reconstruction_stencil: str
    split_reconstruction: Any = None
    frozen_state: str = "ARITHMETIC"


class ConvectiveFluxesSetup(NamedTuple):
    """data_types/numerical_setup/conservatives.py:62-67 (the block of the selected solver is set, like the reference)."""
    convective_solver: str
    godunov: Optional[HighOrderGodunovSetup] = None
    flux_splitting: Optional[FluxSplittingSetup] = None


class DissipativeFluxesSetup(NamedTuple):
    """read_conservatives.py:374-440 (names; CENTRAL4 is what the sm_100a kernels implement)."""
    reconstruction_stencil: str = "CENTRAL4"
    derivative_stencil_center: str = "CENTRAL4"
    derivative_stencil_face: str = "CENTRAL4"
    is_laplacian: bool = False


class PositivitySetup(NamedTuple):
    """read_positivity.py: the interpolation limiter and the SIMPLE / NASA flux limiters are what this path implements."""
    is_interpolation_limiter: bool = False
    limit_velocity: bool = False
    flux_limiter: Optional[str] = None
    flux_partition: str = "UNIFORM"


class ConservativesSetup(NamedTuple):
    halo_cells: int
    time_integration: TimeIntegrationSetup
    convective_fluxes: ConvectiveFluxesSetup
    dissipative_fluxes: DissipativeFluxesSetup = DissipativeFluxesSetup()
    positivity: PositivitySetup = PositivitySetup()


class ActivePhysicsSetup(NamedTuple):
    is_convective_flux: bool = False
    is_viscous_flux: bool = False
    is_heat_flux: bool = False
    is_volume_force: bool = False
    is_surface_tension: bool = False
    is_geometric_source: bool = False
    is_viscous_heat_production: bool = True      # read_active_physics default


class PrecisionSetup(NamedTuple):
    is_double_precision_compute: bool = True
    is_double_precision_output: bool = True


class LoggingSetup(NamedTuple):
    level: str = "INFO"
    frequency: int = 1
    is_positivity: bool = True
    is_only_last_stage: bool = True


class OutputSetup(NamedTuple):
    logging: LoggingSetup = LoggingSetup()


class NumericalSetup(NamedTuple):
    conservatives: ConservativesSetup
    active_physics: ActivePhysicsSetup
    precision: PrecisionSetup
    output: OutputSetup


class GeneralSetup(NamedTuple):
    case_name: str
    end_time: float
    end_step: int
    save_path: str
    save_dt: Any


class DomainSetup(NamedTuple):
    cells: Tuple[int, int, int]
    range: Tuple[Tuple[float, float], ...]
    decomposition: Tuple[int, int, int]


class TransportSetup(NamedTuple):
    """material_properties/transport (read_material_manager.py:200-330): the CUSTOM float values and the
    PRANDTL conductivity model are implemented; non-dimensionalisation references are 1."""
    dynamic_viscosity_model: str = "CUSTOM"
    dynamic_viscosity: float = 0.0
    bulk_viscosity: float = 0.0
    thermal_conductivity_model: str = "CUSTOM"
    thermal_conductivity: float = 0.0
    prandtl_number: float = 1.0


class MaterialSetup(NamedTuple):
    model: str
    specific_heat_ratio: float
    specific_gas_constant: float
    transport: TransportSetup = TransportSetup()


class CaseSetup(NamedTuple):
    general_setup: GeneralSetup
    domain_setup: DomainSetup
    boundary_condition_setup: Dict[str, str]
    initial_condition_setup: Dict[str, Any]
    material_setup: MaterialSetup
    wall_velocity_setup: Dict[str, Tuple[Any, Any, Any]] = {}            # WALL faces: (u, v, w), floats or lambda strings
    # DIRICHLET / NEUMANN / SIMPLE_INFLOW / SIMPLE_OUTFLOW faces: (rho, u, v, w, p), each a float, a lambda string of
    # (active transverse coordinates, t), or None where the type reads no such entry
    dirichlet_setup: Dict[str, Tuple[Any, Any, Any, Any, Any]] = {}
    gravity: Tuple[float, float, float] = (0.0, 0.0, 0.0)               # forcings/gravity
    # faces given as a LIST of types with bounding_domain lambdas (halos/outer/material.py:121-277): face -> the
    # DIRICHLET entries [(values, bounding_domain string)] and the bounding_domain of the entry the kernels fill;
    # boundary_condition_setup[face] names that entry's type (ZEROGRADIENT / SYMMETRY)
    multi_type_setup: Dict[str, Dict[str, Any]] = {}


def _np_namespace():
    """`jnp` as seen by the lambda strings of a case file (setup_reader.py:171): NumPy."""
    return np


def evaluate_dirichlet_face(values, face: str, domain_information, rank: int = 0,
                            callable_name: str = "primitives_callable"):
    """primitives_callable of one DIRICHLET face on this block (halos/outer/material.py:770-790 with
    boundary_condition.py:105-126): floats stay floats; lambda strings are evaluated on the mesh grid (indexing "ij") of
    the ACTIVE transverse cell centres and returned shaped like the face's halo slab with extent 1 along the normal and
    the inactive axes.  Labels must be the active transverse axis names + "t" (read_boundary_conditions.py:156).
    Time-dependent callables are not implemented on the B200 path (the halo data is set up once)."""
    di = domain_information
    ax = FACES.index(face) // 2
    trans = [i for i in di.active_axes_indices if i != ax]
    centers = di.get_device_cell_centers(rank)
    mesh = np.meshgrid(*[np.asarray(centers[i], dtype=np.float64) for i in trans], indexing="ij") if trans else []
    shape = [di.device_number_of_cells[i] if i in trans else 1 for i in range(3)]
    labels = tuple(AXES[i] for i in trans) + ("t",)
    out = []
    for k, v in zip(("rho", "u", "v", "w", "p"), values):
        if not isinstance(v, str):
            out.append(None if v is None else float(v))
            continue
        path = f"boundary_conditions/{face}/{callable_name}/{k}"
        fn = eval(v, {"jnp": _np_namespace(), "np": np})   # noqa: S307 -- same contract as the reference
        names = fn.__code__.co_varnames[:fn.__code__.co_argcount]
        _assert(tuple(names) == labels, f"Input argument labels of lambda for {path} must be {labels}.", "case")
        mshape = mesh[0].shape if mesh else ()
        a = np.broadcast_to(np.asarray(fn(*mesh, 0.0), dtype=np.float64), mshape).reshape(shape)
        b = np.broadcast_to(np.asarray(fn(*mesh, 0.7310585786300049), dtype=np.float64), mshape).reshape(shape)
        if not np.array_equal(a, b, equal_nan=True):
            raise NotImplementedError(f"{path}: a time-dependent DIRICHLET callable is not implemented on the B200 path")
        out.append(np.ascontiguousarray(a))
    return tuple(out)


def evaluate_bounding_domain(expr: str, face: str, domain_information, rank: int = 0) -> np.ndarray:
    """bounding_domain of one entry of a multi-type face (halos/outer/material.py:252-258): a lambda of the ACTIVE
    transverse coordinates, evaluated on this block's transverse cell centres; bool, shaped like the face's halo slab
    with extent 1 along the normal and the inactive axes."""
    di = domain_information
    ax = FACES.index(face) // 2
    trans = [i for i in di.active_axes_indices if i != ax]
    centers = di.get_device_cell_centers(rank)
    mesh = np.meshgrid(*[np.asarray(centers[i], dtype=np.float64) for i in trans], indexing="ij") if trans else []
    shape = [di.device_number_of_cells[i] if i in trans else 1 for i in range(3)]
    fn = eval(expr, {"jnp": _np_namespace(), "np": np})   # noqa: S307 -- same contract as the reference
    names = fn.__code__.co_varnames[:fn.__code__.co_argcount]
    labels = tuple(AXES[i] for i in trans)
    _assert(tuple(names) == labels, f"Input argument labels of lambda for boundary_conditions/{face}/bounding_domain must "
                                    f"be {labels}.", "case")
    mshape = mesh[0].shape if mesh else ()
    return np.ascontiguousarray(np.broadcast_to(np.asarray(fn(*mesh)), mshape).astype(bool).reshape(shape))


def make_ic_callable(value, labels: Tuple[str, ...], path: str) -> Callable:
    """initial_condition entry -> callable on the mesh grid (setup_reader.py:125-218).
    Floats broadcast over the grid; strings are Python lambdas eval'd with `jnp` (and `np`)
    in scope, whose argument names must equal the active axis names."""
    if isinstance(value, bool) or not isinstance(value, (float, int, str)) and not callable(value):
        _assert(False, f"Value of {path} must be float or string that specifies a lambda function.", "case")
    def full_shape(args):
        return np.broadcast_shapes(*[np.shape(a) for a in args])
    if isinstance(value, (float, int)):
        v = float(value)
        return lambda *args: np.broadcast_to(np.float64(v), full_shape(args))
    if isinstance(value, str):
        fn = eval(value, {"jnp": _np_namespace(), "np": np})   # noqa: S307 -- same contract as the reference
        names = fn.__code__.co_varnames[:fn.__code__.co_argcount]
        _assert(tuple(names) == tuple(labels), f"Input argument labels of lambda for {path} must be {labels}.", "case")
        # the mesh may be passed sparse (broadcastable 1-D axes): same values as the reference's dense
        # meshgrid evaluation, without materialising three full-size coordinate arrays
        return lambda *args: np.broadcast_to(np.asarray(fn(*args), dtype=np.float64), full_shape(args))
    return value


class EquationInformation:
    """equation_information.py:92-110 for SINGLE-PHASE."""
    equation_type = "SINGLE-PHASE"
    no_primes = 5
    primes_tuple = ("rho", "u", "v", "w", "p")
    cons_tuple = ("rho", "rhou", "rhov", "rhow", "E")
    ids_mass = 0
    ids_velocity = (1, 2, 3)
    ids_energy = 4
    velocity_minor_axes = ((2, 3), (3, 1), (1, 2))
    levelset_model = False
    diffuse_interface_model = False
    is_compute_temperature = False


class InputManager:
    def __init__(self, case_setup: Union[str, Dict], numerical_setup: Union[str, Dict], materials_setup=None) -> None:
        self.case_setup_dict = read_json_setup(case_setup, "case setup")
        self.numerical_setup_dict = read_json_setup(numerical_setup, "numerical setup")
        if materials_setup:
            raise NotImplementedError("materials_setup")          # input_manager.py:52-55 raises too
        self._check_nondim(self.case_setup_dict)
        self.numerical_setup = self._read_numerical(self.numerical_setup_dict)
        self.case_setup = self._read_case(self.case_setup_dict)
        self.equation_information = EquationInformation()
        ap = self.numerical_setup.active_physics          # equation_information.py: temperature buffer with these on
        self.equation_information.is_compute_temperature = bool(ap.is_viscous_flux or ap.is_heat_flux)
        self.domain_information = DomainInformation(
            cells=self.case_setup.domain_setup.cells,
            domain_range=self.case_setup.domain_setup.range,
            split=self.case_setup.domain_setup.decomposition,
            nh=self.numerical_setup.conservatives.halo_cells)
        self._sanity_check()

    # -- numerical setup ---------------------------------------------------
    def _read_numerical(self, d: Dict) -> NumericalSetup:
        cons_d = get_setup_value(d, "conservatives", "conservatives", dict, False)
        nh = get_setup_value(cons_d, "halo_cells", "conservatives/halo_cells", int, False,
                             numerical_value_condition=(">", 0))
        ti_d = get_setup_value(cons_d, "time_integration", "conservatives/time_integration", dict, False)
        integ = get_setup_value(ti_d, "integrator", "conservatives/time_integration/integrator", str, False)
        integ = R.select(integ, R.REFERENCE_TIME_INTEGRATORS, R.DICT_TIME_INTEGRATION,
                         "conservatives/time_integration/integrator")
        cfl = get_setup_value(ti_d, "CFL", "conservatives/time_integration/CFL", float, True, 0.5,
                              numerical_value_condition=(">", 0.0))
        fixed = get_setup_value(ti_d, "fixed_timestep", "conservatives/time_integration/fixed_timestep", float, True,
                                False, numerical_value_condition=(">", 0.0))
        cf_d = get_setup_value(cons_d, "convective_fluxes", "conservatives/convective_fluxes", dict, False)
        solver = get_setup_value(cf_d, "convective_solver", "conservatives/convective_fluxes/convective_solver", str,
                                 True, "GODUNOV")
        solver = R.select(solver, R.REFERENCE_CONVECTIVE_SOLVERS, R.DICT_CONVECTIVE_SOLVER,
                          "conservatives/convective_fluxes/convective_solver")
        if solver == "FLUX-SPLITTING":
            return InputManager._read_flux_splitting_setup(d, cons_d, cf_d, nh, integ, cfl, fixed)
        base = "conservatives/convective_fluxes/godunov"
        g_d = get_setup_value(cf_d, "godunov", base, dict, False)
        riemann = get_setup_value(g_d, "riemann_solver", base + "/riemann_solver", str, False)
        riemann = R.select(riemann, R.REFERENCE_RIEMANN_SOLVERS, R.DICT_RIEMANN_SOLVER, base + "/riemann_solver")
        sig = get_setup_value(g_d, "signal_speed", base + "/signal_speed", str, False)
        sig = R.select(sig, R.REFERENCE_SIGNAL_SPEEDS, R.DICT_SIGNAL_SPEEDS, base + "/signal_speed")
        rv = get_setup_value(g_d, "reconstruction_variable", base + "/reconstruction_variable", str, False)
        rv = R.select(rv, R.REFERENCE_RECONSTRUCTION_VARIABLES, R.TUPLE_RECONSTRUCTION_VARIABLES,
                      base + "/reconstruction_variable")
        stencil = get_setup_value(g_d, "reconstruction_stencil", base + "/reconstruction_stencil", str, False)
        stencil = R.select(stencil, R.REFERENCE_RECONSTRUCTION_STENCILS, R.DICT_SPATIAL_RECONSTRUCTION,
                           base + "/reconstruction_stencil")
        frozen = get_setup_value(g_d, "frozen_state", base + "/frozen_state", str, True, "ARITHMETIC")
        frozen = R.select(frozen, R.REFERENCE_FROZEN_STATES, R.TUPLE_FROZEN_STATE, base + "/frozen_state")
        InputManager._check_halos(stencil, nh)
        dissipative, positivity, active_physics, precision, logging_setup = InputManager._read_common_blocks(d, cons_d)
        return NumericalSetup(
            ConservativesSetup(nh, TimeIntegrationSetup(integ, cfl, fixed),
                               ConvectiveFluxesSetup(solver, HighOrderGodunovSetup(riemann, sig, stencil, rv, frozen)),
                               dissipative, positivity),
            active_physics, precision, OutputSetup(logging_setup))

    @staticmethod
    def _check_halos(stencil, nh):
        # read_conservatives.py:361-365
        req = R.REQUIRED_HALOS[stencil]
        _assert(nh >= req, f"Reconstruction stencil {stencil} requires at least {req} halo cells, "
                           f"but only {nh} are specified.", "numerical")
        if nh < R.KERNEL_HALOS:
            raise NotImplementedError(f"conservatives/halo_cells = {nh} is valid for {stencil} in JAX-Fluids, but the "
                                      f"B200 sweep kernels stage {R.KERNEL_HALOS} cells on either side of a face: "
                                      f"halo_cells >= {R.KERNEL_HALOS} is required on the B200 path")

    @staticmethod
    def _read_flux_splitting_setup(d, cons_d, cf_d, nh, integ, cfl, fixed):
        """read_conservatives.py:205-244 (read_flux_splitting): convective_solver = FLUX-SPLITTING."""
        base = "conservatives/convective_fluxes/flux_splitting"
        fs_d = get_setup_value(cf_d, "flux_splitting", base, dict, False)
        fs = get_setup_value(fs_d, "flux_splitting", base + "/flux_splitting", str, False)
        fs = R.select(fs, R.REFERENCE_FLUX_SPLITTING, R.TUPLE_FLUX_SPLITTING, base + "/flux_splitting")
        stencil = get_setup_value(fs_d, "reconstruction_stencil", base + "/reconstruction_stencil", str, False)
        stencil = R.select(stencil, R.REFERENCE_RECONSTRUCTION_STENCILS + ("SPLIT-RECONSTRUCTION",),
                           R.DICT_SPATIAL_RECONSTRUCTION, base + "/reconstruction_stencil")
        frozen = get_setup_value(fs_d, "frozen_state", base + "/frozen_state", str, True, "ARITHMETIC")
        frozen = R.select(frozen, R.REFERENCE_FROZEN_STATES, R.TUPLE_FROZEN_STATE, base + "/frozen_state")
        InputManager._check_halos(stencil, nh)
        dissipative, positivity, active_physics, precision, logging_setup = InputManager._read_common_blocks(d, cons_d)
        if positivity.flux_limiter:
            raise NotImplementedError("conservatives/positivity/flux_limiter with convective_solver = FLUX-SPLITTING is "
                                      "not implemented on the B200 path")
        return NumericalSetup(
            ConservativesSetup(nh, TimeIntegrationSetup(integ, cfl, fixed),
                               ConvectiveFluxesSetup("FLUX-SPLITTING", None, FluxSplittingSetup(fs, stencil, None, frozen)),
                               dissipative, positivity),
            active_physics, precision, OutputSetup(logging_setup))

    @staticmethod
    def _read_common_blocks(d, cons_d):
        """positivity, active_physics, dissipative_fluxes, precision, output/logging -- the blocks every convective
        solver shares."""
        pos_d = cons_d.get("positivity", {}) or {}
        for k, v in pos_d.items():
            if k.startswith("is_") and v and k != "is_interpolation_limiter":
                raise NotImplementedError(f"conservatives/positivity/{k} is not implemented on the B200 path "
                                          "(implemented: is_interpolation_limiter)")
        # read_positivity.py:20-30 (flux_limiter: one of TUPLE_POSITIVITY_FIXES or absent / false)
        flux_limiter = pos_d.get("flux_limiter", False)
        if flux_limiter not in (None, False):
            flux_limiter = R.select(flux_limiter, R.REFERENCE_POSITIVITY_FIXES, R.TUPLE_POSITIVITY_FIXES,
                                    "conservatives/positivity/flux_limiter")
        else:
            flux_limiter = None
        flux_partition = R.select(get_setup_value(pos_d, "flux_partition", "conservatives/positivity/flux_partition",
                                                  str, True, "UNIFORM"),
                                  R.REFERENCE_POSITIVITY_PARTITIONS, R.TUPLE_POSITIVITY_PARTITIONS,
                                  "conservatives/positivity/flux_partition")
        positivity = PositivitySetup(
            bool(get_setup_value(pos_d, "is_interpolation_limiter", "conservatives/positivity/is_interpolation_limiter",
                                 bool, True, False)),
            bool(get_setup_value(pos_d, "limit_velocity", "conservatives/positivity/limit_velocity", bool, True, False)),
            flux_limiter, flux_partition)

        ap_d = get_setup_value(d, "active_physics", "active_physics", dict, False)
        ap = {}
        for f in ActivePhysicsSetup._fields:
            default = ActivePhysicsSetup._field_defaults[f]
            ap[f] = bool(get_setup_value(ap_d, f, f"active_physics/{f}", bool, True, default))
        active_physics = ActivePhysicsSetup(**ap)
        _assert(active_physics.is_convective_flux or active_physics.is_viscous_flux or active_physics.is_heat_flux,
                "active_physics: at least one of is_convective_flux, is_viscous_flux, is_heat_flux must be true.",
                "numerical")
        for f in ("is_surface_tension", "is_geometric_source"):
            if getattr(active_physics, f):
                raise NotImplementedError(f"active_physics/{f} is not implemented on the B200 path "
                                          "(single-phase convective + viscous + heat flux + gravity only)")
        # read_conservatives.py:374-440
        is_diss = active_physics.is_viscous_flux or active_physics.is_heat_flux
        df_d = get_setup_value(cons_d, "dissipative_fluxes", "conservatives/dissipative_fluxes", dict, not is_diss, {})
        df = {}
        for key, optional in (("derivative_stencil_face", not is_diss),
                              ("reconstruction_stencil", not active_physics.is_viscous_flux),
                              ("derivative_stencil_center", not active_physics.is_viscous_flux)):
            v = get_setup_value(df_d, key, f"conservatives/dissipative_fluxes/{key}", str, optional, "CENTRAL4")
            if is_diss:
                v = R.select(v, R.REFERENCE_CENTRAL_STENCILS, R.TUPLE_DISSIPATIVE_STENCILS,
                             f"conservatives/dissipative_fluxes/{key}")
            df[key] = v
        df["is_laplacian"] = bool(get_setup_value(df_d, "is_laplacian", "conservatives/dissipative_fluxes/is_laplacian",
                                                  bool, True, False))
        if is_diss and df["is_laplacian"]:
            raise NotImplementedError("conservatives/dissipative_fluxes/is_laplacian is not implemented on the B200 path")
        dissipative = DissipativeFluxesSetup(**df)
        for k in ("active_forcings",):
            for kk, v in (d.get(k, {}) or {}).items():
                if v:
                    raise NotImplementedError(f"{k}/{kk} is not implemented on the B200 path")
        for k in ("levelset", "diffuse_interface"):
            if (d.get(k, {}) or {}).get("model"):
                raise NotImplementedError(f"{k}/model is not implemented on the B200 path (single-phase only)")

        pr_d = d.get("precision", {}) or {}
        precision = PrecisionSetup(
            bool(get_setup_value(pr_d, "is_double_precision_compute", "precision/is_double_precision_compute", bool, True, True)),
            bool(get_setup_value(pr_d, "is_double_precision_output", "precision/is_double_precision_output", bool, True, True)))
        if not precision.is_double_precision_compute:
            raise NotImplementedError("precision/is_double_precision_compute = false is not implemented on the B200 "
                                      "path (fp64 only)")
        out_d = d.get("output", {}) or {}
        log_d = out_d.get("logging", {}) or {}
        logging_setup = LoggingSetup(
            level=get_setup_value(log_d, "level", "output/logging/level", str, True, "INFO",
                                  possible_string_values=("DEBUG", "INFO", "DEBUG_TO_FILE", "INFO_TO_FILE", "NONE")),
            frequency=get_setup_value(log_d, "frequency", "output/logging/frequency", int, True, 1,
                                      numerical_value_condition=(">", 0)),
            is_positivity=bool(get_setup_value(log_d, "is_positivity", "output/logging/is_positivity", bool, True, True)),
            is_only_last_stage=bool(get_setup_value(log_d, "is_only_last_stage", "output/logging/is_only_last_stage",
                                                    bool, True, True)))
        return dissipative, positivity, active_physics, precision, logging_setup

    # -- case setup ----------------------------------------------------------
    @staticmethod
    def _check_nondim(d: Dict):
        nd = d.get("nondimensionalization_parameters", {}) or {}
        for k, v in nd.items():
            if float(v) != 1.0:
                raise NotImplementedError(f"nondimensionalization_parameters/{k} != 1.0 is not implemented on the "
                                          "B200 path")

    def _read_case(self, d: Dict) -> CaseSetup:
        S = "case"
        gen_d = get_setup_value(d, "general", "general", dict, False, setup=S)
        case_name = get_setup_value(gen_d, "case_name", "general/case_name", str, False, setup=S)
        end_time = get_setup_value(gen_d, "end_time", "general/end_time", float, True, False, None, (">=", 0.0), S)
        end_step = get_setup_value(gen_d, "end_step", "general/end_step", int, True, False, None, (">=", 0), S)
        _assert(isinstance(end_step, int) and not isinstance(end_step, bool) or isinstance(end_time, float),
                "Either end_time or end_step must be given.", S)
        if isinstance(end_time, bool):
            end_time = float(np.finfo(np.float64).max)        # read_general.py:45-48
        if isinstance(end_step, bool):
            end_step = int(np.iinfo(np.int64).max)
        save_path = get_setup_value(gen_d, "save_path", "general/save_path", str, True, "./results", setup=S)
        save_dt = get_setup_value(gen_d, "save_dt", "general/save_dt", float, True, False, None, (">", 0.0), S)
        general = GeneralSetup(case_name, end_time, end_step, save_path, save_dt)

        if (d.get("restart", {}) or {}).get("is_restart"):
            raise NotImplementedError("restart/is_restart is not implemented on the B200 path (needs h5py)")

        dom_d = get_setup_value(d, "domain", "domain", dict, False, setup=S)
        cells, rng = [], []
        for ax in AXES:
            a_d = get_setup_value(dom_d, ax, f"domain/{ax}", dict, False, setup=S)
            n = get_setup_value(a_d, "cells", f"domain/{ax}/cells", int, False, None, None, (">", 0), S)
            r = get_setup_value(a_d, "range", f"domain/{ax}/range", list, False, setup=S)
            _assert(len(r) == 2 and r[1] > r[0], f"domain/{ax}/range must be [lower, upper] with upper > lower.", S)
            st = a_d.get("stretching")
            if st and (st.get("type") not in (None, False, "HOMOGENEOUS", "HOMOGENOUS")):
                raise NotImplementedError(f"domain/{ax}/stretching is not implemented on the B200 path")
            cells.append(int(n))
            rng.append((float(r[0]), float(r[1])))
        dec_d = dom_d.get("decomposition", {}) or {}
        split = tuple(int(get_setup_value(dec_d, f"split_{ax}", f"domain/decomposition/split_{ax}", int, True, 1,
                                          None, (">", 0), S)) for ax in AXES)
        for i, ax in enumerate(AXES):
            _assert(cells[i] % split[i] == 0, f"domain/{ax}/cells must be divisible by split_{ax}.", S)
            _assert(not (cells[i] == 1 and split[i] > 1), f"Inactive axis {ax} cannot be split.", S)
        domain = DomainSetup(tuple(cells), tuple(rng), split)

        bc_d = get_setup_value(d, "boundary_conditions", "boundary_conditions", dict, False, setup=S)
        bcs = {}
        walls = {}
        dirichlets = {}
        multi = {}
        for f in FACES:
            f_d = get_setup_value(bc_d, f, f"boundary_conditions/{f}", (dict, list), False, setup=S)
            if isinstance(f_d, list):
                # several types on one face, each with a bounding_domain lambda of the transverse coordinates: one entry
                # of a type the halo kernels fill (ZEROGRADIENT / SYMMETRY) + DIRICHLET entries the host applies
                natives = [e for e in f_d if e.get("type") in ("ZEROGRADIENT", "SYMMETRY")]
                others = [e for e in f_d if e.get("type") not in ("ZEROGRADIENT", "SYMMETRY")]
                for e in f_d:
                    R.select(get_setup_value(e, "type", f"boundary_conditions/{f}/type", str, False, setup=S),
                             R.REFERENCE_BOUNDARY_TYPES, R.TUPLE_BOUNDARY_TYPES, f"boundary_conditions/{f}/type", S)
                    get_setup_value(e, "bounding_domain", f"boundary_conditions/{f}/bounding_domain", str, False, setup=S)
                if len(natives) != 1 or any(e["type"] != "DIRICHLET" for e in others):
                    raise NotImplementedError(
                        f"boundary_conditions/{f}: several types per face are implemented on the B200 path for one "
                        "ZEROGRADIENT or SYMMETRY entry combined with DIRICHLET entries")
                entries = []
                for e in others:
                    pc_d = get_setup_value(e, "primitives_callable", f"boundary_conditions/{f}/primitives_callable", dict,
                                           False, setup=S)
                    vals = tuple(v if isinstance(v, str) else float(v) for v in (
                        get_setup_value(pc_d, k, f"boundary_conditions/{f}/primitives_callable/{k}", (float, str), False,
                                        setup=S) for k in ("rho", "u", "v", "w", "p")))
                    entries.append((vals, e["bounding_domain"]))
                multi[f] = {"dirichlet": entries, "kernel_bounding_domain": natives[0]["bounding_domain"]}
                f_d = {"type": natives[0]["type"]}
            t = get_setup_value(f_d, "type", f"boundary_conditions/{f}/type", str, False, setup=S)
            t = R.select(t, R.REFERENCE_BOUNDARY_TYPES, R.TUPLE_BOUNDARY_TYPES, f"boundary_conditions/{f}/type", S)
            active = cells[FACE_AXIS[f]] > 1
            _assert((t == "INACTIVE") == (not active),
                    f"boundary_conditions/{f}/type must be INACTIVE exactly for inactive axes.", S)
            bcs[f] = t
            if t == "WALL":
                # read_boundary_conditions: wall_velocity_callable {u, v, w}, floats or lambdas of (coords, t)
                wv_d = get_setup_value(f_d, "wall_velocity_callable", f"boundary_conditions/{f}/wall_velocity_callable",
                                       dict, False, setup=S)
                uvw = []
                for k in ("u", "v", "w"):
                    v = get_setup_value(wv_d, k, f"boundary_conditions/{f}/wall_velocity_callable/{k}", (float, str),
                                        False, setup=S)
                    # a lambda of the face's active transverse coordinates and the time, like primitives_callable;
                    # evaluated on the block's transverse cell centres by the runtime (time-independent only)
                    uvw.append(v if isinstance(v, str) else float(v))
                walls[f] = tuple(uvw)
            if t in R.BOUNDARY_VALUE_KEYS:
                # read_boundary_conditions: primitives_callable {rho, u, v, w, p} (SIMPLE_INFLOW: no p, SIMPLE_OUTFLOW:
                # p only), floats or lambdas of (coords, t); None for the entries the type does not read
                pc_d = get_setup_value(f_d, "primitives_callable", f"boundary_conditions/{f}/primitives_callable",
                                       dict, False, setup=S)
                vals = []
                for k in ("rho", "u", "v", "w", "p"):
                    if k not in R.BOUNDARY_VALUE_KEYS[t]:
                        vals.append(None)
                        continue
                    v = get_setup_value(pc_d, k, f"boundary_conditions/{f}/primitives_callable/{k}", (float, str),
                                        False, setup=S)
                    # a lambda of the face's active transverse coordinates and the time (read_boundary_conditions.py:156);
                    # evaluated on the block's transverse cell centres by evaluate_dirichlet_face (time-independent only)
                    vals.append(v if isinstance(v, str) else float(v))
                dirichlets[f] = tuple(vals)
        for ax, (hi, lo) in enumerate((("east", "west"), ("north", "south"), ("top", "bottom"))):
            _assert((bcs[hi] == "PERIODIC") == (bcs[lo] == "PERIODIC"),
                    f"PERIODIC boundary conditions must be set at both {hi} and {lo}.", S)

        ic_d = get_setup_value(d, "initial_condition", "initial_condition", dict, False, setup=S)
        if "primitives" in ic_d:
            ic_d = ic_d["primitives"]
        labels = tuple(ax for i, ax in enumerate(AXES) if cells[i] > 1)
        ic = {}
        if "turbulent" in ic_d:
            # initial_condition/turbulent (read_initial_conditions.py:95-170): a generated turbulent field instead of
            # lambdas; the HIT generator (turbulence/initialization/hit.py) is what this path implements
            tb = get_setup_value(ic_d, "turbulent", "initial_condition/turbulent", dict, False, setup=S)
            tcase = get_setup_value(tb, "case", "initial_condition/turbulent/case", str, False, setup=S)
            if tcase != "HIT":
                raise NotImplementedError(f"initial_condition/turbulent/case '{tcase}' is not implemented on the B200 "
                                          "path (implemented: HIT)")
            seed = get_setup_value(tb, "random_seed", "initial_condition/turbulent/random_seed", int, True, 0,
                                   numerical_value_condition=(">=", 0), setup=S)
            pp = "initial_condition/turbulent/parameters"
            par = get_setup_value(tb, "parameters", pp, dict, False, setup=S)
            ic["turbulent"] = dict(
                case=tcase, random_seed=seed,
                T_ref=get_setup_value(par, "T_ref", f"{pp}/T_ref", float, False, numerical_value_condition=(">", 0.0), setup=S),
                rho_ref=get_setup_value(par, "rho_ref", f"{pp}/rho_ref", float, False, numerical_value_condition=(">", 0.0), setup=S),
                ma_target=get_setup_value(par, "ma_target", f"{pp}/ma_target", float, False, numerical_value_condition=(">", 0.0), setup=S),
                energy_spectrum=get_setup_value(par, "energy_spectrum", f"{pp}/energy_spectrum", str, False,
                                                possible_string_values=("KOLMOGOROV", "EXPONENTIAL", "BOX"), setup=S),
                ic_type=get_setup_value(par, "ic_type", f"{pp}/ic_type", str, False,
                                        possible_string_values=("IC1", "IC2", "IC3", "IC4"), setup=S),
                xi_0=get_setup_value(par, "xi_0", f"{pp}/xi_0", int, False, numerical_value_condition=(">=", 0), setup=S),
                xi_1=get_setup_value(par, "xi_1", f"{pp}/xi_1", int, True, 16, numerical_value_condition=(">=", 0), setup=S),
                is_velocity_spectral=get_setup_value(par, "is_velocity_spectral", f"{pp}/is_velocity_spectral", bool, True,
                                                     False, setup=S))
            _assert(cells[0] == cells[1] == cells[2] and cells[0] > 1,
                    "initial_condition/turbulent/case HIT needs a cubic 3-D grid.", S)
        else:
            for name in ("rho", "u", "v", "w", "p"):
                _assert(name in ic_d, f"Key {name} is not optional, but missing initial_condition/{name}.", S)
                ic[name] = make_ic_callable(ic_d[name], labels, f"initial_condition/{name}")

        mp_d = get_setup_value(d, "material_properties", "material_properties", dict, False, setup=S)
        eos_d = get_setup_value(mp_d, "equation_of_state", "material_properties/equation_of_state", dict, False, setup=S)
        model = get_setup_value(eos_d, "model", "material_properties/equation_of_state/model", str, False, setup=S)
        model = R.select(model, R.REFERENCE_MATERIALS, R.DICT_MATERIAL, "material_properties/equation_of_state/model", S)
        gamma = get_setup_value(eos_d, "specific_heat_ratio", "material_properties/equation_of_state/specific_heat_ratio",
                                float, False, None, None, (">", 1.0), S)
        Rgas = get_setup_value(eos_d, "specific_gas_constant",
                               "material_properties/equation_of_state/specific_gas_constant", float, True, 1.0, None,
                               (">", 0.0), S)
        # forcings: gravity (read_forcings; used with active_physics/is_volume_force) is what this path implements
        gravity = (0.0, 0.0, 0.0)
        fo_d = d.get("forcings", {}) or {}
        for k, v in fo_d.items():
            if k != "gravity" and v:
                raise NotImplementedError(f"forcings/{k} is not implemented on the B200 path (implemented: gravity)")
        if self.numerical_setup.active_physics.is_volume_force:
            gv = get_setup_value(fo_d, "gravity", "forcings/gravity", list, False, setup=S)
            _assert(len(gv) == 3 and all(isinstance(x, (int, float)) for x in gv),
                    "forcings/gravity must be a list of three numbers.", S)
            gravity = tuple(float(x) for x in gv)
        transport = self._read_transport(mp_d)
        return CaseSetup(general, domain, bcs, ic, MaterialSetup(model, float(gamma), float(Rgas), transport), walls,
                         dirichlets, gravity, multi)

    def _read_transport(self, mp_d: Dict) -> TransportSetup:
        """read_material_manager.py:200-330: required exactly when the flux that needs them is active."""
        S = "case"
        ap = self.numerical_setup.active_physics
        base = "material_properties/transport"
        if not (ap.is_viscous_flux or ap.is_heat_flux):
            return TransportSetup()
        tr_d = get_setup_value(mp_d, "transport", base, dict, False, setup=S)

        def custom_float(dd, key, path):
            v = get_setup_value(dd, key, path, (float, str), False, setup=S)
            if isinstance(v, str):
                raise NotImplementedError(f"{path} given as a lambda string is not implemented on the B200 path "
                                          "(constant CUSTOM values only)")
            return float(v)
        mu_model, mu, bulk = "CUSTOM", 0.0, 0.0
        need_mu = ap.is_viscous_flux or (tr_d.get("thermal_conductivity", {}) or {}).get("model") == "PRANDTL"
        if need_mu:
            mu_d = get_setup_value(tr_d, "dynamic_viscosity", base + "/dynamic_viscosity", dict, False, setup=S)
            mu_model = get_setup_value(mu_d, "model", base + "/dynamic_viscosity/model", str, False,
                                       possible_string_values=("CUSTOM", "SUTHERLAND"), setup=S)
            if mu_model != "CUSTOM":
                raise NotImplementedError(f"{base}/dynamic_viscosity/model = '{mu_model}' is a valid JAX-Fluids option "
                                          "that is not implemented on the B200 path (implemented: ('CUSTOM',))")
            mu = custom_float(mu_d, "value", base + "/dynamic_viscosity/value")
            bulk = float(get_setup_value(tr_d, "bulk_viscosity", base + "/bulk_viscosity", float, False, setup=S))
        tc_model, tc, prandtl = "CUSTOM", 0.0, 1.0
        if ap.is_heat_flux:
            tc_d = get_setup_value(tr_d, "thermal_conductivity", base + "/thermal_conductivity", dict, False, setup=S)
            tc_model = get_setup_value(tc_d, "model", base + "/thermal_conductivity/model", str, False,
                                       possible_string_values=("CUSTOM", "PRANDTL", "SUTHERLAND"), setup=S)
            if tc_model == "CUSTOM":
                tc = custom_float(tc_d, "value", base + "/thermal_conductivity/value")
            elif tc_model == "PRANDTL":
                prandtl = float(get_setup_value(tc_d, "prandtl_number", base + "/thermal_conductivity/prandtl_number",
                                                float, False, numerical_value_condition=(">", 0.0), setup=S))
            else:
                raise NotImplementedError(f"{base}/thermal_conductivity/model = '{tc_model}' is a valid JAX-Fluids "
                                          "option that is not implemented on the B200 path "
                                          "(implemented: ('CUSTOM', 'PRANDTL'))")
        return TransportSetup(mu_model, mu, bulk, tc_model, tc, prandtl)

    def _sanity_check(self):
        di = self.domain_information
        nh = self.numerical_setup.conservatives.halo_cells
        for i in di.active_axes_indices:
            _assert(di.device_number_of_cells[i] >= nh,
                    f"domain/{AXES[i]}/cells per device must be >= halo_cells.", "case")

    # -- helpers used by the other managers -----------------------------------
    @property
    def gamma(self) -> float:
        return self.case_setup.material_setup.specific_heat_ratio
