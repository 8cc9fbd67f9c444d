"""Option names of the reference and which of them this path implements.

The JSON strings of the reference select the new path unchanged.  A name the
reference knows but this path does not implement raises NotImplementedError
("... not implemented on the B200 path"); a name the reference does not know
fails the same consistency assertion the reference raises.

Name lists: solvers/convective_fluxes/__init__.py:7-12, solvers/riemann_solvers/
__init__.py:16-34, stencils/reconstruction/**/__init__.py, time_integration/
__init__.py:6-11, materials/single_materials/__init__.py:8-15, halos/outer/__init__.py:1-7.
"""
REFERENCE_CONVECTIVE_SOLVERS = ("GODUNOV", "FLUX-SPLITTING", "ALDM", "CENTRAL")
REFERENCE_RIEMANN_SOLVERS = ("LAX-FRIEDRICHS", "HLL", "HLLC", "HLLC_SIMPLEALPHA", "HLLC-LM", "RUSANOV", "AUSMP", "CATUM")
REFERENCE_SIGNAL_SPEEDS = ("ARITHMETIC", "RUSANOV", "DAVIS", "DAVIS2", "EINFELDT", "TORO")
REFERENCE_RECONSTRUCTION_VARIABLES = ("PRIMITIVE", "CONSERVATIVE", "CHAR-PRIMITIVE", "CHAR-CONSERVATIVE")
REFERENCE_FROZEN_STATES = ("ARITHMETIC", "ROE")
REFERENCE_POSITIVITY_FIXES = ("SIMPLE", "NASA", "HAS")               # solvers/positivity/__init__.py:1-3
REFERENCE_POSITIVITY_PARTITIONS = ("UNIFORM", "CELLSIZE", "WAVESPEED")   # :5-7
REFERENCE_RECONSTRUCTION_STENCILS = (
    "KOREN", "MC", "MINMOD", "SUPERBEE", "VANALBADA", "VANLEER", "KOREN-ADAP", "MC-ADAP", "MINMOD-ADAP",
    "SUPERBEE-ADAP", "VANALBADA-ADAP", "VANLEER-ADAP", "MINMOD-AD", "MINMOD-AD-ADAP", "TENO5", "TENO5-A",
    "TENO5-ADAP", "TENO6", "TENO6-ADAP", "TENO6-A", "TENO6-A-ADAP", "TENO8", "TENO8-A", "WENO1", "WENO3-JS",
    "WENO3-N", "WENO3-NN-OPT1", "WENO3-NN-OPT2", "WENO3-Z", "WENO3-FP", "WENO5-JS", "WENO5-Z", "WENO6-CU",
    "WENO6-CUM1", "WENO6-CUM2", "WENO7-JS", "WENO9-JS", "WENO3-JS-ADAP", "WENO3-Z-ADAP", "WENO5-JS-ADAP",
    "WENO5-Z-ADAP", "WENO6-CU-ADAP", "CENTRAL2", "CENTRAL2-ADAP", "CENTRAL4", "CENTRAL4-ADAP", "CENTRAL6",
    "CENTRAL6-ADAP", "CENTRAL8", "CENTRAL8-ADAP")
REFERENCE_CENTRAL_STENCILS = ("CENTRAL2", "CENTRAL4", "CENTRAL6", "CENTRAL8", "CENTRAL2-ADAP", "CENTRAL4-ADAP",
                              "CENTRAL6-ADAP", "CENTRAL8-ADAP")
REFERENCE_TIME_INTEGRATORS = ("EULER", "RK2", "RK3", "RK2_LS4")
REFERENCE_MATERIALS = ("IdealGas", "SafeIdealGas", "StiffenedGas", "StiffenedGasComplete", "Tait",
                       "BarotropicCavitationFluid")
REFERENCE_BOUNDARY_TYPES = (
    "ZEROGRADIENT", "SYMMETRY", "PERIODIC", "INACTIVE", "LINEAREXTRAPOLATION", "WALL", "ISOTHERMALWALL",
    "MASSTRANSFERWALL", "MASSTRANSFERWALL_PARAMETERIZED", "ISOTHERMALMASSTRANSFERWALL", "DIRICHLET", "NEUMANN",
    "SIMPLE_INFLOW", "SIMPLE_OUTFLOW", "DIRICHLET_PARAMETERIZED", "OPPOSITIONCONTROLWALL",
    "ISOTHERMALOPPOSITIONCONTROLWALL")

# what the sm_100a kernels implement
DICT_CONVECTIVE_SOLVER = {"GODUNOV": "HighOrderGodunov", "FLUX-SPLITTING": "FluxSplittingScheme"}
REFERENCE_FLUX_SPLITTING = ("ROE", "CLLF", "LLF", "CLF")     # solvers/__init__.py:1-3
# "CLF" passes the reference's input check but no branch of eigendecomposition.py:664-703 handles it (it raises there)
TUPLE_FLUX_SPLITTING = ("ROE", "CLLF", "LLF")
# HLLC is the tuned kernel; the others ride on the RUSANOV kernel instantiations (run-time variants).  Not
# implemented: LAX-FRIEDRICHS (a global max over all faces before every sweep), CATUM, HLLC_SIMPLEALPHA (its
# single-phase branch raises NotImplementedError in the reference itself)
DICT_RIEMANN_SOLVER = {"HLLC": "HLLC", "RUSANOV": "Rusanov", "HLL": "HLL", "HLLC-LM": "HLLCLM", "AUSMP": "AUSMP"}
DICT_SIGNAL_SPEEDS = {"EINFELDT": "signal_speed_Einfeldt", "ARITHMETIC": "signal_speed_Arithmetic",
                      "RUSANOV": "signal_speed_Rusanov", "DAVIS": "signal_speed_Davis", "TORO": "signal_speed_Toro"}
# WENO5-Z / WENO5-JS are the tuned kernels; the others run in the generic (reference-order) instantiations
DICT_SPATIAL_RECONSTRUCTION = {"WENO5-Z": "WENO5Z", "WENO5-JS": "WENO5JS", "WENO1": "WENO1", "WENO3-JS": "WENO3JS",
                               "WENO3-Z": "WENO3Z", "TENO5": "TENO5", "WENO6-CU": "WENO6CU", "KOREN": "KOREN",
                               "MC": "MC", "MINMOD": "MINMOD", "SUPERBEE": "SUPERBEE", "VANALBADA": "VANALBADA",
                               "VANLEER": "VANLEER", "WENO3-N": "WENO3N",
                               "CENTRAL2": "CentralSecondOrderReconstruction", "TENO6": "TENO6", "TENO5-A": "TENO5A",
                               "TENO6-A": "TENO6A"}
# PRIMITIVE / CHAR-PRIMITIVE with the ARITHMETIC frozen state are the tuned kernels; the conservative forms and ROE run
# in the generic (reference-order) instantiations
TUPLE_RECONSTRUCTION_VARIABLES = ("PRIMITIVE", "CONSERVATIVE", "CHAR-PRIMITIVE", "CHAR-CONSERVATIVE")
TUPLE_FROZEN_STATE = ("ARITHMETIC", "ROE")
TUPLE_POSITIVITY_FIXES = ("SIMPLE", "NASA")          # HAS is marked "TODO NEEDS UPDATE" upstream (limiter_flux.py:211)
TUPLE_POSITIVITY_PARTITIONS = ("UNIFORM", "CELLSIZE")   # WAVESPEED needs a global max per axis before every sweep
TUPLE_DISSIPATIVE_STENCILS = ("CENTRAL4",)     # reconstruction / derivative_center / derivative_face
DICT_TIME_INTEGRATION = {"EULER": "Euler", "RK2": "RungeKutta2", "RK3": "RungeKutta3", "RK2_LS4": "RungeKutta2_LS4"}
DICT_MATERIAL = {"IdealGas": "IdealGas"}
# NEUMANN / SIMPLE_INFLOW / SIMPLE_OUTFLOW: the kernels' base rule on these faces is ZEROGRADIENT; the prescribed data is
# applied on top of it in the kernels from per-face device arrays (runtime.BlockRuntime._make_face_data)
TUPLE_BOUNDARY_TYPES = ("ZEROGRADIENT", "SYMMETRY", "PERIODIC", "INACTIVE", "WALL", "DIRICHLET", "NEUMANN",
                        "SIMPLE_INFLOW", "SIMPLE_OUTFLOW")
# entries of primitives_callable each of these types reads (read_boundary_conditions.py:160-365)
BOUNDARY_VALUE_KEYS = {"DIRICHLET": ("rho", "u", "v", "w", "p"), "NEUMANN": ("rho", "u", "v", "w", "p"),
                       "SIMPLE_INFLOW": ("rho", "u", "v", "w"), "SIMPLE_OUTFLOW": ("p",)}

# required_halos of the reference classes (weno5_base.py:18, weno3_base.py:18, weno6_base.py:17, muscl3.py:20,
# weno1_js.py: 1, central_2.py:21, teno6_base.py:16)
REQUIRED_HALOS = {"WENO5-Z": 3, "WENO5-JS": 3, "WENO1": 1, "WENO3-JS": 2, "WENO3-Z": 2, "TENO5": 3, "WENO6-CU": 3,
                  "KOREN": 2, "MC": 2, "MINMOD": 2, "SUPERBEE": 2, "VANALBADA": 2, "VANLEER": 2, "WENO3-N": 2,
                  "CENTRAL2": 1, "TENO6": 3, "TENO5-A": 3, "TENO6-A": 3}
KERNEL_HALOS = 3          # the sweep kernels always stage 3 cells on either side of a face


def select(value, reference_names, implemented, path, setup="numerical"):
    """Reference-style validation + 'not implemented on the B200 path'."""
    assert isinstance(value, str), (
        f"Consistency error in {setup} setup file. Key {path} must be of types {str}, but is of type {type(value)}.")
    assert value in reference_names, (
        f"Consistency error in {setup} setup file. Value of {path} must be in {tuple(reference_names)} "
        "if value is of type str.")
    if value not in implemented:
        raise NotImplementedError(
            f"{path} = '{value}' is a valid JAX-Fluids option that is not implemented on the B200 path "
            f"(implemented: {tuple(implemented)}).")
    return value
