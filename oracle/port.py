"""CPU oracle: NumPy restatement of JAX-Fluids' single-phase convective hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in `jaxfluids_b200/` may import this module;
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs use it, and only as the checker / CPU baseline.

Pinning: this restatement performs the reference's elementwise operations in
the reference's order, so on NumPy it is BIT-IDENTICAL to the reference's own
Python sources executed on the NumPy `jax` stand-in (oracle/refharness).  That
identity is asserted by tests/test_oracle_pinning.py (in the build container,
where /root/reference exists) and, everywhere, against the committed fixtures
tests/golden/*.npz generated from the reference by oracle/refharness/make_goldens.py.
The reference ships no tests/golden vectors of its own for this path.

All file:line citations are relative to /root/reference/src/jaxfluids/.

Layout: buffers are C-order (5, X, Y, Z); an active axis carries `nh` halo
cells on both sides, an inactive axis has extent 1 (initialization/
helper_functions.py:49).  Variables: prims (rho,u,v,w,p), cons (rho,rho u,rho v,
rho w,E) (equation_information.py:92-110).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Sequence, Tuple

import numpy as np

EPS = float(np.finfo(np.float64).eps)          # config/precision.py:44-55 (eps, fp64)
STENCIL_EPS = 1e-30                            # config/precision.py:53 (spatial stencil eps, fp64)
FACES = ("east", "west", "north", "south", "top", "bottom")   # domain/__init__.py:5-7
FACE_AXIS = {"east": 0, "west": 0, "north": 1, "south": 1, "top": 2, "bottom": 2}
MINOR_AXES = ((2, 3), (3, 1), (1, 2))          # equation_information.py:110 (velocity_minor_axes)

RK = {  # time_integration/{euler,RK2,RK3,RK2_LS4}.py
    "EULER": dict(stages=1, dt_mult=(1.0,), blend=()),
    # RK2_LS4.py:27-30 (the blend table there has three (0, 1) entries: every later stage restarts from U^n)
    "RK2_LS4": dict(stages=4, dt_mult=(0.11, 0.2766, 0.5, 1.0), blend=((0.0, 1.0), (0.0, 1.0), (0.0, 1.0))),
    "RK2": dict(stages=2, dt_mult=(1.0, 0.5), blend=((0.5, 0.5),)),
    "RK3": dict(stages=3, dt_mult=(1.0, 0.25, 2.0 / 3.0), blend=((0.25, 0.75), (2.0 / 3.0, 1.0 / 3.0))),
}


@dataclass
class Setup:
    """The subset of case + numerical setup the path depends on."""
    cells: Tuple[int, int, int]
    domain: Tuple[Tuple[float, float], ...]      # ((xlo,xhi),(ylo,yhi),(zlo,zhi))
    bc: Dict[str, str]                           # face -> PERIODIC|SYMMETRY|ZEROGRADIENT|INACTIVE
    gamma: float
    nh: int = 5
    recon: str = "CHAR-PRIMITIVE"                # or PRIMITIVE
    stencil: str = "WENO5-Z"                     # or WENO5-JS (godunov.reconstruction_stencil)
    riemann: str = "HLLC"                        # or RUSANOV
    signal_speed: str = "EINFELDT"               # HLLC wave-speed estimate: EINFELDT | ARITHMETIC | RUSANOV | DAVIS | TORO
    integrator: str = "RK3"
    frozen_state: str = "ARITHMETIC"             # godunov/frozen_state or flux_splitting/frozen_state: ARITHMETIC | ROE
    convective_solver: str = "GODUNOV"           # or FLUX-SPLITTING (convective_fluxes/flux_splitting block)
    flux_splitting: str = "ROE"                  # flux_splitting/flux_splitting: ROE | CLLF | LLF (eigenvalue choice)
    cfl: float = 0.5
    fixed_timestep: float | None = None
    inv_dx_override: Tuple[float, float, float] | None = None   # sub-blocks of a larger grid (port_mt, multi-block tests)
    # dissipative fluxes (active_physics + material_properties/transport); CENTRAL4 stencils
    is_viscous_flux: bool = False
    is_heat_flux: bool = False
    is_viscous_heat_production: bool = True      # numerical_setup/active_physics default
    dynamic_viscosity: float = 0.0               # transport/dynamic_viscosity, model CUSTOM (float)
    bulk_viscosity: float = 0.0
    thermal_conductivity_model: str = "CUSTOM"   # CUSTOM (float value) | PRANDTL
    thermal_conductivity: float = 0.0
    prandtl_number: float = 1.0
    gas_constant: float = 1.0                    # equation_of_state/specific_gas_constant
    # conservatives/positivity (limiter_interpolation.py) and WALL boundaries (halos/outer/material.py:473-520)
    is_interpolation_limiter: bool = False
    limit_velocity: bool = False
    flux_limiter: str | None = None              # positivity/flux_limiter: SIMPLE | NASA (limiter_flux.py)
    flux_partition: str = "UNIFORM"              # positivity/flux_partition: UNIFORM | CELLSIZE
    wall_velocity: Dict[str, Tuple[float, float, float]] = field(default_factory=dict)   # face -> (u, v, w), constants
    # face -> (rho, u, v, w, p): constants, or arrays over the transverse interior cells of the face shaped like the
    # halo slab with extent 1 along the face normal (a lambda of the transverse coordinates in the case file)
    dirichlet: Dict[str, Tuple] = field(default_factory=dict)
    # NEUMANN / SIMPLE_INFLOW / SIMPLE_OUTFLOW faces -> (rho, u, v, w, p) of their primitives_callable, None where the
    # type takes no entry (SIMPLE_INFLOW: no p; SIMPLE_OUTFLOW: p only); floats or arrays as for `dirichlet`
    bc_values: Dict[str, Tuple] = field(default_factory=dict)
    # faces with several types (a list in the case file): face -> [dict(kind, mask, values)], mask = bounding_domain on the
    # face's transverse cells (bool, shaped like the halo slab with extent 1 along the normal), values for DIRICHLET
    # entries; bc[face] then names the entry the sm_100a kernels fill (the others are applied by the host runtime)
    bc_multi: Dict[str, list] = field(default_factory=dict)
    # active_physics/is_volume_force + forcings/gravity (source_term_solver.py:163-186)
    is_volume_force: bool = False
    gravity: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    is_convective_flux: bool = True              # active_physics/is_convective_flux (false: heat-equation examples)
    active: Tuple[int, ...] = field(init=False)

    def __post_init__(self):
        self.cells = tuple(int(c) for c in self.cells)
        self.active = tuple(i for i in range(3) if self.cells[i] > 1)

    # domain/mesh_creation/homogenous.py:13 ; domain_information.py:290
    @property
    def dx(self):
        return tuple(np.float64((self.domain[i][1] - self.domain[i][0]) / self.cells[i]) for i in range(3))

    @property
    def inv_dx(self):
        if self.inv_dx_override is not None:
            return tuple(np.float64(v) for v in self.inv_dx_override)
        return tuple(np.float64(1.0) / d for d in self.dx)

    @property
    def dx_min(self):                            # domain_information.py:697-702
        return min(self.dx[i] for i in self.active)

    @property
    def shape(self):                             # initialization/helper_functions.py:49
        return (5,) + tuple(n + 2 * self.nh if n > 1 else 1 for n in self.cells)

    @property
    def interior(self):
        nh = self.nh
        return tuple(slice(nh, -nh) if self.cells[i] > 1 else slice(None) for i in range(3))

    @property
    def is_dissipative(self):
        return self.is_viscous_flux or self.is_heat_flux

    @property
    def cp(self):                                # ideal_gas.py:33
        return self.gamma / (self.gamma - 1.0) * self.gas_constant

    def cell_centers(self):                      # homogenous.py:14
        out = []
        for i in range(3):
            lo, hi = self.domain[i]
            d = (hi - lo) / self.cells[i]
            out.append(np.linspace(lo + d / 2, hi - d / 2, self.cells[i]))
        return out


# --------------------------------------------------------------------------
# equation of state / variable transforms
# --------------------------------------------------------------------------
def cons_from_prims(p, gamma):
    """equation_manager.py:93-101 ; ideal_gas.py:85-88 ; math/sum_consistent.py:22-24."""
    rho = p[0]
    e = p[4] / (rho * (gamma - 1.0))
    E = rho * (0.5 * (np.square(p[1]) + np.square(p[2]) + np.square(p[3])) + e)
    return np.stack([rho, rho * p[1], rho * p[2], rho * p[3], E], axis=0)


def prims_from_cons(c, gamma):
    """equation_manager.py:164-171 ; ideal_gas.py:73-75."""
    rho = c[0]
    one_rho = 1.0 / rho
    u, v, w = c[1] * one_rho, c[2] * one_rho, c[3] * one_rho
    e = c[4] * one_rho - 0.5 * (np.square(u) + np.square(v) + np.square(w))
    p = (gamma - 1.0) * e * rho
    return np.stack([rho, u, v, w, p], axis=0)


def speed_of_sound(p, rho, gamma):
    """ideal_gas.py:69-71."""
    return np.sqrt(gamma * p / rho)


def physical_flux(p, c, axis):
    """equation_manager.py:237-252."""
    m = c[axis + 1]
    f = [m, m * p[1], m * p[2], m * p[3], p[axis + 1] * (c[4] + p[4])]
    f[axis + 1] = f[axis + 1] + p[4]
    return np.stack(f, axis=0)


# --------------------------------------------------------------------------
# halo fill (outer boundaries)
# --------------------------------------------------------------------------
def halo_fill(prims, cons, s: Setup):
    """halos/halo_manager.py:146-234 -> halos/outer/material.py:94-287, 868-894.

    Source slices: halos/outer/boundary_condition.py:563-595; symmetry sign of the
    face-normal velocity :698-731; faces in the order east..bottom; transverse
    range = interior; mask = 1.0 so `halo*(1-mask) + new*mask`.
    Returns new (prims, cons)."""
    prims, cons = prims.copy(), cons.copy()
    nh = s.nh
    inter = s.interior
    for face in FACES:
        ax = FACE_AXIS[face]
        if ax not in s.active or s.bc[face] == "INACTIVE":
            continue
        hi = face in ("east", "north", "top")
        # several types on one face (case file: a list with bounding_domain lambdas, material.py:121-277): applied one
        # after the other, each through its mask over the face's transverse cells
        entries = s.bc_multi.get(face) or [dict(kind=s.bc[face], mask=1.0, values=None)]
        for ent in entries:
            kind, mask = ent["kind"], ent["mask"]
            if kind == "PERIODIC":
                src = slice(nh, 2 * nh) if hi else slice(-2 * nh, -nh)
            elif kind in ("SYMMETRY", "WALL"):
                src = slice(-nh - 1, -2 * nh - 1, -1) if hi else slice(2 * nh - 1, nh - 1, -1)
            elif kind in ("ZEROGRADIENT", "DIRICHLET", "NEUMANN", "SIMPLE_INFLOW", "SIMPLE_OUTFLOW"):
                src = slice(-nh - 1, -nh) if hi else slice(nh, nh + 1)     # boundary_condition.py:580-595
            else:
                raise NotImplementedError(kind)
            dst = slice(-nh, None) if hi else slice(0, nh)
            sl_src = [slice(None)] + list(inter)
            sl_dst = [slice(None)] + list(inter)
            sl_src[1 + ax] = src
            sl_dst[1 + ax] = dst
            hp = prims[tuple(sl_src)]
            if kind == "DIRICHLET":                  # halos/outer/material.py:732-798: primitives_callable -- constants, or
                vals = ent["values"] if ent["values"] is not None else s.dirichlet[face]   # arrays over the transverse cells
                hp = np.stack([np.ones_like(hp[0]) * vals[v] for v in range(5)], axis=0)
            if kind == "NEUMANN":                    # :825-866: last interior cell + (value * upwind sign) * dx
                vals = s.bc_values[face]
                sgn = -1 if hi else 1                # material.py:41-44
                hp = np.stack([hp[v] + (np.ones_like(hp[0]) * vals[v]) * sgn * s.dx[ax] for v in range(5)], axis=0)
            if kind == "SIMPLE_INFLOW":              # :966-1022: density and velocity prescribed, pressure from inside
                vals = s.bc_values[face]
                hp = np.stack([np.ones_like(hp[0]) * vals[v] for v in range(4)] + [hp[4]], axis=0)
            if kind == "SIMPLE_OUTFLOW":             # :1024-1050: everything from inside, pressure prescribed
                vals = s.bc_values[face]
                hp = np.stack([hp[v] for v in range(4)] + [np.ones_like(hp[0]) * vals[4]], axis=0)
            if kind == "SYMMETRY":
                sign = np.ones((5, 1, 1, 1))
                sign[1 + ax] *= -1.0
                hp = hp * sign
            if kind == "WALL":                       # halos/outer/material.py:473-520: u_halo = 2 u_wall - u_mirror
                uw = s.wall_velocity.get(face, (0.0, 0.0, 0.0))
                hp = np.stack([hp[0]] + [2 * (np.ones_like(hp[0]) * uw[k]) - hp[1 + k] for k in range(3)] + [hp[4]], axis=0)
            hc = cons_from_prims(hp, s.gamma)
            prims[tuple(sl_dst)] = prims[tuple(sl_dst)] * (1 - mask) + hp * mask       # material.py:270-275
            cons[tuple(sl_dst)] = cons[tuple(sl_dst)] * (1 - mask) + hc * mask
    if s.is_dissipative and len(s.active) > 1:       # halo_manager.py:119-129, :193-199
        prims, cons = edge_halo_fill(prims, cons, s)
    return prims, cons


# domain/__init__.py:9-15 (order matters only for documentation: edge regions are disjoint and no edge
# reads another edge's region)
EDGES = ("west_south", "west_north", "east_north", "east_south", "south_bottom", "north_bottom",
         "south_top", "north_top", "east_bottom", "west_bottom", "east_top", "west_top")
OPPOSITE = {"east": "west", "west": "east", "north": "south", "south": "north", "top": "bottom", "bottom": "top"}


def _edge_range(face, code, nh):
    """halos/halo_slices.py:98-147: index range along the axis of `face` for the edge region itself
    (code None), or with this face's range moved to the nh interior layers next to it (code '1')."""
    hi = face in ("east", "north", "top")
    if code == "1":
        return slice(-2 * nh, -nh) if hi else slice(nh, 2 * nh)
    return slice(-nh, None) if hi else slice(0, nh)


def edge_halo_fill(prims, cons, s: Setup):
    """halos/outer/material.py:289-383 (edge_halo_update / compute_edge_halos) with the type combination of
    boundary_condition.py:128-179 and the retrieve table :607-655: the FIRST face of the edge that is
    PERIODIC or SYMMETRY decides (PERIODIC: copy from the opposite side, SYMMETRY: mirror + negate that
    face's normal velocity); otherwise the mean of the two adjacent halo regions.  cons recomputed."""
    prims, cons = prims.copy(), cons.copy()
    nh = s.nh
    for edge in EDGES:
        fa, fb = edge.split("_")
        axa, axb = FACE_AXIS[fa], FACE_AXIS[fb]
        if axa not in s.active or axb not in s.active:
            continue
        types = []
        decided = False
        for f in (fa, fb):
            if not decided and s.bc[f] in ("PERIODIC", "SYMMETRY"):
                types.append(s.bc[f])
                decided = True
            else:
                types.append("ANY")

        def region(face_a, code_a, face_b, code_b):
            sl = [slice(None)] + list(s.interior)
            sl[1 + axa] = _edge_range(face_a, code_a, nh)
            sl[1 + axb] = _edge_range(face_b, code_b, nh)
            return tuple(sl)

        if types == ["ANY", "ANY"]:
            hp = 0.5 * (prims[region(fa, "1", fb, None)] + prims[region(fa, None, fb, "1")])
        else:
            k = 0 if types[0] != "ANY" else 1
            faces = [fa, fb]
            codes = [None, None]
            codes[k] = "1"
            if types[k] == "PERIODIC":
                faces[k] = OPPOSITE[faces[k]]
            hp = prims[region(faces[0], codes[0], faces[1], codes[1])]
            if types[k] == "SYMMETRY":
                ax = (axa, axb)[k]
                flip = [slice(None)] * 4
                flip[1 + ax] = slice(None, None, -1)
                hp = hp[tuple(flip)]
                sign = np.ones((5, 1, 1, 1))
                sign[1 + ax] *= -1.0
                hp = hp * sign
        hc = cons_from_prims(hp, s.gamma)
        dst = region(fa, None, fb, None)
        prims[dst] = prims[dst] * (1 - 1) + hp * 1
        cons[dst] = cons[dst] * (1 - 1) + hc * 1
    return prims, cons


# --------------------------------------------------------------------------
# WENO5-Z
# --------------------------------------------------------------------------
_DR = (1 / 10, 6 / 10, 3 / 10)
_CR = ((1 / 3, -7 / 6, 11 / 6), (-1 / 6, 5 / 6, 1 / 3), (1 / 3, 5 / 6, -1 / 6))


def weno5z(a, b, c, d, e):
    """weno5_base.py:34-51 and weno/weno5_z.py:32-52 on the 5 cells (i-2..i+2) (j=0)
    or their mirror (i+3..i-1) (j=1)."""
    beta_0 = 13.0 / 12.0 * np.square(a - 2 * b + c) + 1.0 / 4.0 * np.square(a - 4 * b + 3 * c)
    beta_1 = 13.0 / 12.0 * np.square(b - 2 * c + d) + 1.0 / 4.0 * np.square(b - d)
    beta_2 = 13.0 / 12.0 * np.square(c - 2 * d + e) + 1.0 / 4.0 * np.square(3 * c - 4 * d + e)
    tau_5 = np.abs(beta_0 - beta_2)
    alpha_0 = _DR[0] * (1.0 + tau_5 / (beta_0 + STENCIL_EPS))
    alpha_1 = _DR[1] * (1.0 + tau_5 / (beta_1 + STENCIL_EPS))
    alpha_2 = _DR[2] * (1.0 + tau_5 / (beta_2 + STENCIL_EPS))
    one_alpha = 1.0 / (alpha_0 + alpha_1 + alpha_2)
    omega_0, omega_1, omega_2 = alpha_0 * one_alpha, alpha_1 * one_alpha, alpha_2 * one_alpha
    p_0 = _CR[0][0] * a + _CR[0][1] * b + _CR[0][2] * c
    p_1 = _CR[1][0] * b + _CR[1][1] * c + _CR[1][2] * d
    p_2 = _CR[2][0] * c + _CR[2][1] * d + _CR[2][2] * e
    return omega_0 * p_0 + omega_1 * p_1 + omega_2 * p_2


def weno5js(a, b, c, d, e):
    """weno5_base.py:34-51 and weno/weno5_js.py:32-50."""
    beta_0 = 13.0 / 12.0 * np.square(a - 2 * b + c) + 1.0 / 4.0 * np.square(a - 4 * b + 3 * c)
    beta_1 = 13.0 / 12.0 * np.square(b - 2 * c + d) + 1.0 / 4.0 * np.square(b - d)
    beta_2 = 13.0 / 12.0 * np.square(c - 2 * d + e) + 1.0 / 4.0 * np.square(3 * c - 4 * d + e)
    one_beta_0_sq = 1.0 / (beta_0 * beta_0 + STENCIL_EPS)
    one_beta_1_sq = 1.0 / (beta_1 * beta_1 + STENCIL_EPS)
    one_beta_2_sq = 1.0 / (beta_2 * beta_2 + STENCIL_EPS)
    alpha_0 = _DR[0] * one_beta_0_sq
    alpha_1 = _DR[1] * one_beta_1_sq
    alpha_2 = _DR[2] * one_beta_2_sq
    one_alpha = 1.0 / (alpha_0 + alpha_1 + alpha_2)
    omega_0, omega_1, omega_2 = alpha_0 * one_alpha, alpha_1 * one_alpha, alpha_2 * one_alpha
    p_0 = _CR[0][0] * a + _CR[0][1] * b + _CR[0][2] * c
    p_1 = _CR[1][0] * b + _CR[1][1] * c + _CR[1][2] * d
    p_2 = _CR[2][0] * c + _CR[2][1] * d + _CR[2][2] * e
    return omega_0 * p_0 + omega_1 * p_1 + omega_2 * p_2


def weno1(q, j):
    """weno/weno1_js.py:24-29: the upwind cell."""
    return q[2]


_DR3 = (1 / 3, 2 / 3)
_CR3 = ((-0.5, 1.5), (0.5, 0.5))


def _weno3_parts(q):
    """weno3_base.py:33-49 on (u_im, u_i, u_ip) = the upwind-biased cells i-1, i, i+1."""
    u_im, u_i, u_ip = q[1], q[2], q[3]
    beta_0 = np.square(u_i - u_im)
    beta_1 = np.square(u_ip - u_i)
    p_0 = _CR3[0][0] * u_im + _CR3[0][1] * u_i
    p_1 = _CR3[1][0] * u_i + _CR3[1][1] * u_ip
    return beta_0, beta_1, p_0, p_1


def weno3js(q, j):
    """weno/weno3_js.py:15-43."""
    beta_0, beta_1, p_0, p_1 = _weno3_parts(q)
    one_beta_0_sq = 1.0 / (beta_0 * beta_0 + STENCIL_EPS)
    one_beta_1_sq = 1.0 / (beta_1 * beta_1 + STENCIL_EPS)
    alpha_0 = _DR3[0] * one_beta_0_sq
    alpha_1 = _DR3[1] * one_beta_1_sq
    one_alpha = 1.0 / (alpha_0 + alpha_1)
    return (alpha_0 * one_alpha) * p_0 + (alpha_1 * one_alpha) * p_1


def weno3z(q, j):
    """weno/weno3_z.py:23-43."""
    beta_0, beta_1, p_0, p_1 = _weno3_parts(q)
    tau_3 = np.abs(beta_0 - beta_1)
    alpha_z_0 = _DR3[0] * (1.0 + tau_3 / (beta_0 + STENCIL_EPS))
    alpha_z_1 = _DR3[1] * (1.0 + tau_3 / (beta_1 + STENCIL_EPS))
    one_alpha_z = 1.0 / (alpha_z_0 + alpha_z_1)
    return (alpha_z_0 * one_alpha_z) * p_0 + (alpha_z_1 * one_alpha_z) * p_1


def weno3n(q, j):
    """weno/weno3_n.py:22-41: WENO3-Z weights with tau_3 = |(beta_0 + beta_1)/2 - beta_3|."""
    beta_0, beta_1, p_0, p_1 = _weno3_parts(q)
    u_im, u_i, u_ip = q[1], q[2], q[3]
    beta_3 = 13 / 12 * np.square(u_im - 2 * u_i + u_ip) + 1 / 4 * np.square(u_im - u_ip)
    tau_3 = np.abs(0.5 * (beta_0 + beta_1) - beta_3)
    alpha_z_0 = _DR3[0] * (1.0 + tau_3 / (beta_0 + STENCIL_EPS))
    alpha_z_1 = _DR3[1] * (1.0 + tau_3 / (beta_1 + STENCIL_EPS))
    one_alpha_z = 1.0 / (alpha_z_0 + alpha_z_1)
    return (alpha_z_0 * one_alpha_z) * p_0 + (alpha_z_1 * one_alpha_z) * p_1


def central2(q, j):
    """reconstruction/central/central_2.py:38-47 (the one central stencil stencils/__init__.py:19 offers the convective
    reconstruction): the mean of the two cells of the face, on both sides."""
    return 0.5 * (q[2] + q[3])


def _weno5_parts(a, b, c, d, e):
    """weno5_base.py:34-51."""
    beta_0 = 13.0 / 12.0 * np.square(a - 2 * b + c) + 1.0 / 4.0 * np.square(a - 4 * b + 3 * c)
    beta_1 = 13.0 / 12.0 * np.square(b - 2 * c + d) + 1.0 / 4.0 * np.square(b - d)
    beta_2 = 13.0 / 12.0 * np.square(c - 2 * d + e) + 1.0 / 4.0 * np.square(3 * c - 4 * d + e)
    p_0 = _CR[0][0] * a + _CR[0][1] * b + _CR[0][2] * c
    p_1 = _CR[1][0] * b + _CR[1][1] * c + _CR[1][2] * d
    p_2 = _CR[2][0] * c + _CR[2][1] * d + _CR[2][2] * e
    return beta_0, beta_1, beta_2, p_0, p_1, p_2


_DR_TENO5 = (0.05, 0.55, 0.40)                  # teno/teno5.py:26 ("optimized spectral properties")


def teno5(q, j):
    """teno/teno5.py:32-71: C = 1, q = 6, C_T = 1e-5; sharp cut-off of the sub-stencils."""
    beta_0, beta_1, beta_2, p_0, p_1, p_2 = _weno5_parts(q[0], q[1], q[2], q[3], q[4])
    tau_5 = np.abs(beta_0 - beta_2)
    gamma_0 = np.power(1.0 + tau_5 / (beta_0 + STENCIL_EPS), 6)
    gamma_1 = np.power(1.0 + tau_5 / (beta_1 + STENCIL_EPS), 6)
    gamma_2 = np.power(1.0 + tau_5 / (beta_2 + STENCIL_EPS), 6)
    one_gamma_sum = 1.0 / (gamma_0 + gamma_1 + gamma_2)
    w0 = _DR_TENO5[0] * np.where(gamma_0 * one_gamma_sum < 1e-5, 0, 1)
    w1 = _DR_TENO5[1] * np.where(gamma_1 * one_gamma_sum < 1e-5, 0, 1)
    w2 = _DR_TENO5[2] * np.where(gamma_2 * one_gamma_sum < 1e-5, 0, 1)
    one_dk = 1.0 / (w0 + w1 + w2 + STENCIL_EPS)
    return (w0 * one_dk) * p_0 + (w1 * one_dk) * p_1 + (w2 * one_dk) * p_2


_DR6 = (1 / 20, 9 / 20, 9 / 20, 1 / 20)
_CR6_3 = (11 / 6, -7 / 6, 1 / 3)


def weno6cu(q, j):
    """weno6_base.py:32-58 and weno/weno6_cu.py:36-63 (C = 20) on the six cells i-2..i+3 (mirrored for j=1)."""
    u_imm, u_im, u_i, u_ip, u_ipp, u_ippp = q
    beta_0, beta_1, beta_2, p_0, p_1, p_2 = _weno5_parts(u_imm, u_im, u_i, u_ip, u_ipp)
    beta_3 = 1.0 / 10080 / 12 * (
        271779 * u_imm * u_imm +
        u_imm * (-2380800 * u_im + 4086352 * u_i - 3462252 * u_ip + 1458762 * u_ipp - 245620 * u_ippp) +
        u_im * (5653317 * u_im - 20427884 * u_i + 17905032 * u_ip - 7727988 * u_ipp + 1325006 * u_ippp) +
        u_i * (19510972 * u_i - 35817664 * u_ip + 15929912 * u_ipp - 2792660 * u_ippp) +
        u_ip * (17195652 * u_ip - 15880404 * u_ipp + 2863984 * u_ippp) +
        u_ipp * (3824847 * u_ipp - 1429976 * u_ippp) +
        139633 * u_ippp * u_ippp)
    p_3 = _CR6_3[0] * u_ip + _CR6_3[1] * u_ipp + _CR6_3[2] * u_ippp
    tau_6 = beta_3 - 1 / 6 * (beta_0 + 4 * beta_1 + beta_2)
    alpha_0 = _DR6[0] * (20 + tau_6 / (beta_0 + STENCIL_EPS))
    alpha_1 = _DR6[1] * (20 + tau_6 / (beta_1 + STENCIL_EPS))
    alpha_2 = _DR6[2] * (20 + tau_6 / (beta_2 + STENCIL_EPS))
    alpha_3 = _DR6[3] * (20 + tau_6 / (beta_3 + STENCIL_EPS))
    one_alpha = 1.0 / (alpha_0 + alpha_1 + alpha_2 + alpha_3)
    return ((alpha_0 * one_alpha) * p_0 + (alpha_1 * one_alpha) * p_1 + (alpha_2 * one_alpha) * p_2 +
            (alpha_3 * one_alpha) * p_3)


_DR_TENO6 = (0.050, 0.450, 0.300, 0.200)        # teno/teno6.py:23 (6th-order convergence)
_CR_TENO6_3 = (3 / 12, 13 / 12, -5 / 12, 1 / 12)


def teno6(q, j):
    """teno6_base.py:32-62 and teno/teno6.py:42-73 (C = 1, q = 6, C_T = 1e-7) on the six cells i-2..i+3."""
    u_imm, u_im, u_i, u_ip, u_ipp, u_ippp = q
    beta_0, beta_1, beta_2, p_0, p_1, p_2 = _weno5_parts(u_imm, u_im, u_i, u_ip, u_ipp)
    beta_3 = 1.0 / 240.0 * (
        u_i * (2107 * u_i - 9402 * u_ip + 7042 * u_ipp - 1854 * u_ippp)
        + u_ip * (11003 * u_ip - 17246 * u_ipp + 4642 * u_ippp)
        + u_ipp * (7043 * u_ipp - 3882 * u_ippp)
        + 547 * u_ippp * u_ippp)
    beta_6 = 1.0 / 10080 / 12 * (
        271779 * u_imm * u_imm +
        u_imm * (-2380800 * u_im + 4086352 * u_i - 3462252 * u_ip + 1458762 * u_ipp - 245620 * u_ippp) +
        u_im * (5653317 * u_im - 20427884 * u_i + 17905032 * u_ip - 7727988 * u_ipp + 1325006 * u_ippp) +
        u_i * (19510972 * u_i - 35817664 * u_ip + 15929912 * u_ipp - 2792660 * u_ippp) +
        u_ip * (17195652 * u_ip - 15880404 * u_ipp + 2863984 * u_ippp) +
        u_ipp * (3824847 * u_ipp - 1429976 * u_ippp) +
        139633 * u_ippp * u_ippp)
    p_3 = _CR_TENO6_3[0] * u_i + _CR_TENO6_3[1] * u_ip + _CR_TENO6_3[2] * u_ipp + _CR_TENO6_3[3] * u_ippp
    tau_6 = np.abs(beta_6 - 1 / 6 * (beta_0 + 4 * beta_1 + beta_2))
    gamma_0 = (1.0 + tau_6 / (beta_0 + STENCIL_EPS)) ** 6
    gamma_1 = (1.0 + tau_6 / (beta_1 + STENCIL_EPS)) ** 6
    gamma_2 = (1.0 + tau_6 / (beta_2 + STENCIL_EPS)) ** 6
    gamma_3 = (1.0 + tau_6 / (beta_3 + STENCIL_EPS)) ** 6
    one_gamma_sum = 1.0 / (gamma_0 + gamma_1 + gamma_2 + gamma_3)
    w0 = _DR_TENO6[0] * np.where(gamma_0 * one_gamma_sum < 1e-7, 0, 1)
    w1 = _DR_TENO6[1] * np.where(gamma_1 * one_gamma_sum < 1e-7, 0, 1)
    w2 = _DR_TENO6[2] * np.where(gamma_2 * one_gamma_sum < 1e-7, 0, 1)
    w3 = _DR_TENO6[3] * np.where(gamma_3 * one_gamma_sum < 1e-7, 0, 1)
    one_dk = 1.0 / (w0 + w1 + w2 + w3 + STENCIL_EPS)
    return (w0 * one_dk) * p_0 + (w1 * one_dk) * p_1 + (w2 * one_dk) * p_2 + (w3 * one_dk) * p_3


def teno5a(q, j):
    """teno/teno5_a.py:46-101.  The class stores its optimised weights in `dr_` but reads `_dr`, i.e. the WENO5 weights
    of its base class (weno5_base.py:21) -- restated as the reference computes."""
    u_imm, u_im, u_i, u_ip, u_ipp = q[:5]
    beta_0, beta_1, beta_2, p_0, p_1, p_2 = _weno5_parts(u_imm, u_im, u_i, u_ip, u_ipp)
    tau_5 = np.abs(beta_0 - beta_2)
    gamma_0 = np.power(1.0 + tau_5 / (beta_0 + STENCIL_EPS), 6)
    gamma_1 = np.power(1.0 + tau_5 / (beta_1 + STENCIL_EPS), 6)
    gamma_2 = np.power(1.0 + tau_5 / (beta_2 + STENCIL_EPS), 6)
    one_gamma_sum = 1.0 / (gamma_0 + gamma_1 + gamma_2)
    # eta_k pairs (u_i - u_im, u_im - u_imm), (u_ip - u_i, u_i - u_im), (u_ipp - u_ip, u_ip - u_i): :62-68, written there
    # with the later difference first
    CT = _teno_a_ct_pairs([(u_i - u_im, u_im - u_imm), (u_ip - u_i, u_i - u_im), (u_ipp - u_ip, u_ip - u_i)], 0.24, 10.0, 5.0,
                          nested_min=True)
    w0 = _DR[0] * np.where(gamma_0 * one_gamma_sum < CT, 0, 1)
    w1 = _DR[1] * np.where(gamma_1 * one_gamma_sum < CT, 0, 1)
    w2 = _DR[2] * np.where(gamma_2 * one_gamma_sum < CT, 0, 1)
    one_dk = 1.0 / (w0 + w1 + w2 + STENCIL_EPS)
    return (w0 * one_dk) * p_0 + (w1 * one_dk) * p_1 + (w2 * one_dk) * p_2


def _teno_a_ct_pairs(pairs, Cr, alpha_1, alpha_2, nested_min=False):
    """C_T from explicit (later, earlier) difference pairs; the arithmetic of both TENO-A classes:
    (|2 a b| + eps) / (a^2 + b^2 + eps), the minimum over the pairs, m, g, beta_bar, 10^-beta_bar."""
    eps_d = 0.9 * Cr / (1 - Cr) * 1e-3 ** 2
    etas = [(np.abs(2 * a * b) + eps_d) / (a ** 2 + b ** 2 + eps_d) for a, b in pairs]
    if nested_min:                                  # teno5_a.py:70: min(eta_0, min(eta_1, eta_2))
        eta = np.minimum(etas[0], np.minimum(etas[1], etas[2]))
    else:                                           # teno6_a.py:124-137: running minimum
        eta = etas[0]
        for e in etas[1:]:
            eta = np.minimum(eta, e)
    m = 1 - np.minimum(1.0, eta / Cr)
    g = np.power((1 - m), 4) * (1 + 4 * m)
    beta_bar = alpha_1 - alpha_2 * (1 - g)
    beta_bar = np.ceil(beta_bar) - 1.0
    return np.power(10, (-beta_bar))


_DR_TENO6A = (0.0855682281039113, 0.4294317718960898, 0.1727270875843552, 0.3122729124156450)


def teno6a(q, j):
    """teno/teno6_a.py:42-102 (C = 1, q = 6, Cr = 0.17, alpha = (10.5, 4.5), |beta_3|, |beta_6|, no eps in 1/sum w)."""
    u_imm, u_im, u_i, u_ip, u_ipp, u_ippp = q
    beta_0, beta_1, beta_2, p_0, p_1, p_2 = _weno5_parts(u_imm, u_im, u_i, u_ip, u_ipp)
    beta_3 = 1.0 / 240.0 * (
        u_i * (2107 * u_i - 9402 * u_ip + 7042 * u_ipp - 1854 * u_ippp)
        + u_ip * (11003 * u_ip - 17246 * u_ipp + 4642 * u_ippp)
        + u_ipp * (7043 * u_ipp - 3882 * u_ippp)
        + 547 * u_ippp * u_ippp)
    beta_6 = 1.0 / 10080 / 12 * (
        271779 * u_imm * u_imm +
        u_imm * (-2380800 * u_im + 4086352 * u_i - 3462252 * u_ip + 1458762 * u_ipp - 245620 * u_ippp) +
        u_im * (5653317 * u_im - 20427884 * u_i + 17905032 * u_ip - 7727988 * u_ipp + 1325006 * u_ippp) +
        u_i * (19510972 * u_i - 35817664 * u_ip + 15929912 * u_ipp - 2792660 * u_ippp) +
        u_ip * (17195652 * u_ip - 15880404 * u_ipp + 2863984 * u_ippp) +
        u_ipp * (3824847 * u_ipp - 1429976 * u_ippp) +
        139633 * u_ippp * u_ippp)
    p_3 = _CR_TENO6_3[0] * u_i + _CR_TENO6_3[1] * u_ip + _CR_TENO6_3[2] * u_ipp + _CR_TENO6_3[3] * u_ippp
    beta_3 = np.abs(beta_3)
    beta_6 = np.abs(beta_6)
    tau_6 = np.abs(beta_6 - 1 / 6 * (beta_0 + 4 * beta_1 + beta_2))
    gamma_0 = (1.0 + tau_6 / (beta_0 + STENCIL_EPS)) ** 6
    gamma_1 = (1.0 + tau_6 / (beta_1 + STENCIL_EPS)) ** 6
    gamma_2 = (1.0 + tau_6 / (beta_2 + STENCIL_EPS)) ** 6
    gamma_3 = (1.0 + tau_6 / (beta_3 + STENCIL_EPS)) ** 6
    one_gamma_sum = 1.0 / (gamma_0 + gamma_1 + gamma_2 + gamma_3)
    d = [u_im - u_imm, u_i - u_im, u_ip - u_i, u_ipp - u_ip, u_ippp - u_ipp]
    CT = _teno_a_ct_pairs([(d[1], d[0]), (d[2], d[1]), (d[3], d[2]), (d[4], d[3])], 0.17, 10.5, 4.5)
    w0 = _DR_TENO6A[0] * np.where(gamma_0 * one_gamma_sum < CT, 0, 1)
    w1 = _DR_TENO6A[1] * np.where(gamma_1 * one_gamma_sum < CT, 0, 1)
    w2 = _DR_TENO6A[2] * np.where(gamma_2 * one_gamma_sum < CT, 0, 1)
    w3 = _DR_TENO6A[3] * np.where(gamma_3 * one_gamma_sum < CT, 0, 1)
    one_dk = 1.0 / (w0 + w1 + w2 + w3)
    return (w0 * one_dk) * p_0 + (w1 * one_dk) * p_1 + (w2 * one_dk) * p_2 + (w3 * one_dk) * p_3


# stencils/limiter.py:6-22
MUSCL_LIMITERS = {
    "KOREN": lambda r: np.maximum(0, np.minimum(2 * r, np.minimum((1 + 2 * r) / 3, 2))),
    "MC": lambda r: np.maximum(0, np.minimum(2 * r, np.minimum((1 + r) / 2, 2))),
    "MINMOD": lambda r: np.maximum(0, np.minimum(1, r)),
    "SUPERBEE": lambda r: np.maximum(0, np.maximum(np.minimum(1, 2 * r), np.minimum(2, r))),
    "VANALBADA": lambda r: np.maximum(0, r) * (1 + r) / (1 + r * r),
    "VANLEER": lambda r: np.maximum(0, 2 * r) / (1 + np.abs(r)),
}


def _muscl3(limiter):
    def fn(q, j):
        """muscl/muscl3.py:39-77: the two sides are NOT mirror images of one formula (sign of the differences)."""
        s0, s1, s2 = q[1], q[2], q[3]
        if j == 0:
            delta_central = s2 - s1
            delta_upwind = s1 - s0
        else:
            delta_central = s1 - s2
            delta_upwind = s0 - s1
        r = np.where(delta_upwind >= STENCIL_EPS, delta_central / (delta_upwind + 1e-10),
                     (delta_central + STENCIL_EPS) / (delta_upwind + STENCIL_EPS))
        lim = MUSCL_LIMITERS[limiter](r)
        if j == 0:
            return s1 + 0.5 * lim * delta_upwind
        return s1 - 0.5 * lim * delta_upwind
    return fn


# name -> f(q, j): q = the six cells around the face in upwind-biased order (q[2] | q[3] is the face;
# j = 0: cells i-2..i+3, j = 1: their mirror i+3..i-2), spatial_stencil.py:45-113
STENCILS = {"WENO5-Z": lambda q, j: weno5z(*q[:5]), "WENO5-JS": lambda q, j: weno5js(*q[:5]),
            "WENO1": weno1, "WENO3-JS": weno3js, "WENO3-Z": weno3z, "TENO5": teno5, "WENO6-CU": weno6cu,
            **{name: _muscl3(name) for name in MUSCL_LIMITERS}, "WENO3-N": weno3n, "CENTRAL2": central2,
            "TENO6": teno6, "TENO5-A": teno5a, "TENO6-A": teno6a}
# halo cells the stencil itself needs (required_halos of the reference classes; the sm_100a kernels always stage
# 3 cells on either side of a face)
REQUIRED_HALOS = {"WENO5-Z": 3, "WENO5-JS": 3, "WENO1": 1, "WENO3-JS": 2, "WENO3-Z": 2, "TENO5": 3, "WENO6-CU": 3,
                  **{name: 2 for name in MUSCL_LIMITERS}, "WENO3-N": 2, "CENTRAL2": 1, "TENO6": 3, "TENO5-A": 3, "TENO6-A": 3}


def _window(prims, axis, s: Setup):
    """The six cells k=i-2..i+3 around every face f (cell i = nh-1+f), transverse interior.
    eigendecomposition.py:75-95,117 ; stencils/spatial_stencil.py:45-113."""
    nh, n = s.nh, s.cells[axis]
    out = []
    for k in range(6):
        sl = [slice(None)] + list(s.interior)
        lo = nh - 3 + k
        sl[1 + axis] = slice(lo, lo + n + 1)
        out.append(prims[tuple(sl)])
    return out


# --------------------------------------------------------------------------
# reconstruction
# --------------------------------------------------------------------------
def reconstruct(prims, axis, s: Setup, cons=None):
    """high_order_godunov.py:233-419 (PRIMITIVE :267-280, CONSERVATIVE :282-296, CHAR-PRIMITIVE :298-316,:401-402,
    CHAR-CONSERVATIVE :404-417).  Returns (prims_L, prims_R, cons_L, cons_R), each (5, faces...).  The two
    conservative forms read the conservative buffer (`cons`; formed from the primitives when not given)."""
    w = _window(prims, axis, s)
    stencil = STENCILS[s.stencil]                # the stencil the JSON names
    cl = cr = None
    if s.recon in ("CONSERVATIVE", "CHAR-CONSERVATIVE"):
        wc = _window(cons_from_prims(prims, s.gamma) if cons is None else cons, axis, s)
        if s.recon == "CONSERVATIVE":
            cl = stencil(wc, 0)
            cr = stencil(wc[::-1], 1)
        else:
            # eigendecomposition.py:576-715 at the frozen state of the face, transformtochar / transformtophysical :717-743
            Rm, Lm, _ = conservative_eigensystem(w[2], w[3], axis, s.gamma, None, s.frozen_state)
            chars = [_matvec(Lm, x) for x in wc]
            cl = _matvec(Rm, stencil(chars, 0))
            cr = _matvec(Rm, stencil(chars[::-1], 1))
        pl, pr = prims_from_cons(cl, s.gamma), prims_from_cons(cr, s.gamma)
    elif s.recon == "PRIMITIVE":
        pl = stencil(w, 0)
        pr = stencil(w[::-1], 1)
    elif s.recon == "CHAR-PRIMITIVE":
        g = s.gamma
        ua = 1 + axis
        m0, m1 = MINOR_AXES[axis]
        # rows the tangential velocities occupy in characteristic space
        # (eigendecomposition.py:69-73): axis0 -> (2,3); axis1 -> (3,2); axis2 -> (2,3)
        e0, e1 = ((2, 3), (3, 2), (2, 3))[axis]
        # frozen state of the face from cells i and i+1 (eigendecomposition.py:139-148, 215-231; ROE :233-276)
        ave, _, _, c_ave, cc_ave, _ = frozen_state(w[2], w[3], g, s.frozen_state)
        rho_ave = ave[0]
        # to characteristic variables (eigendecomposition.py:425-431)
        chars = []
        for x in w:
            o = [None] * 5
            o[0] = -0.5 / c_ave * x[ua] + 0.5 / (cc_ave * rho_ave) * x[4]
            o[1] = x[0] - 1.0 / cc_ave * x[4]
            o[e0] = x[m0]
            o[e1] = x[m1]
            o[4] = 0.5 / c_ave * x[ua] + 0.5 / (cc_ave * rho_ave) * x[4]
            chars.append(np.stack(o, axis=0))
        char_l = stencil(chars, 0)
        char_r = stencil(chars[::-1], 1)
        # back to primitives (eigendecomposition.py:517-521)
        res = []
        for x in (char_l, char_r):
            o = [None] * 5
            o[0] = rho_ave * (x[0] + x[4]) + x[1]
            o[ua] = c_ave * (-x[0] + x[4])
            o[m0] = x[e0]
            o[m1] = x[e1]
            o[4] = cc_ave * rho_ave * (x[0] + x[4])
            res.append(np.stack(o, axis=0))
        pl, pr = res
    else:
        raise NotImplementedError(s.recon)
    if s.is_interpolation_limiter:               # high_order_godunov.py:163-174 (returns the conservatives of the limited state)
        pl = _limit_interpolation(pl, w[2], s)   # WENO1 left state = cell i
        pr = _limit_interpolation(pr, w[3], s)   # WENO1 right state = cell i+1
        cl = cr = None
    if cl is None:
        cl, cr = cons_from_prims(pl, s.gamma), cons_from_prims(pr, s.gamma)
    return pl, pr, cl, cr


INTERPOLATION_LIMITER_EPS = (1e-12, 1e-10)       # config/precision.py:54 (density, pressure), fp64


def _limit_interpolation(p, p_first_order, s: Setup):
    """solvers/positivity/limiter_interpolation.py:77-209, SINGLE-PHASE branch: where the reconstructed density
    is below eps (then where the pressure is), fall back to the first-order state -- density and pressure only,
    or all primitives with positivity/limit_velocity."""
    ids = [0, 1, 2, 3, 4] if s.limit_velocity else [0, 4]

    def apply(mask, p):
        p = p.copy()
        for v in ids:
            p[v] = p[v] * (1 - mask) + p_first_order[v] * mask
        return p
    p = apply(np.where(p[0] < INTERPOLATION_LIMITER_EPS[0], 1, 0), p)
    p = apply(np.where(p[4] + 0.0 < INTERPOLATION_LIMITER_EPS[1], 1, 0), p)
    return p


# --------------------------------------------------------------------------
# Riemann solvers
# --------------------------------------------------------------------------
def einfeldt(u_L, u_R, a_L, a_R, rho_L, rho_R):
    """solvers/riemann_solvers/signal_speeds.py:109-133."""
    sL, sR = np.sqrt(rho_L), np.sqrt(rho_R)
    one_dens = 1.0 / (sL + sR)
    eta2 = 0.5 * sL * sR * one_dens * one_dens
    u_bar = (sL * u_L + sR * u_R) * one_dens
    d_bar = np.sqrt((sL * a_L * a_L + sR * a_R * a_R) * one_dens + eta2 * np.square(u_R - u_L))
    return np.minimum(u_bar - d_bar, u_L - a_L), np.maximum(u_bar + d_bar, u_R + a_R)


def signal_speeds(name, u_L, u_R, a_L, a_R, rho_L, rho_R, p_L, p_R, gamma):
    """solvers/riemann_solvers/signal_speeds.py:10-69, :109-157 (+ estimate_pressure :201-214)."""
    if name == "EINFELDT":
        return einfeldt(u_L, u_R, a_L, a_R, rho_L, rho_R)
    if name == "ARITHMETIC":
        u_mean = 0.5 * (u_L + u_R)
        a_mean = 0.5 * (a_L + a_R)
        return np.minimum(u_mean - a_mean, u_L - a_L), np.maximum(u_mean + a_mean, u_R + a_R)
    if name == "RUSANOV":
        S_plus = np.maximum(np.abs(u_L) + a_L, np.abs(u_R) + a_R)
        return -S_plus, S_plus
    if name == "DAVIS":
        return np.minimum(u_L - a_L, u_R - a_R), np.maximum(u_L + a_L, u_R + a_R)
    if name == "TORO":
        rho_bar = 0.5 * (rho_L + rho_R)
        a_bar = 0.5 * (a_L + a_R)
        p_pvrs = 0.5 * (p_L + p_R) - 0.5 * (u_R - u_L) * rho_bar * a_bar
        p_star = np.maximum(0.0, p_pvrs)
        gamma_ = (gamma + 1) * 0.5 / gamma
        q_L = 1.0 * (p_star <= p_L) + np.sqrt(1 + gamma_ * (p_star / p_L - 1)) * (p_star > p_L)
        q_R = 1.0 * (p_star <= p_R) + np.sqrt(1 + gamma_ * (p_star / p_R - 1)) * (p_star > p_R)
        return u_L - a_L * q_L, u_R + a_R * q_R
    raise NotImplementedError(name)


def sstar(u_L, u_R, p_L, p_R, rho_L, rho_R, S_L, S_R):
    """signal_speeds.py:159-199."""
    dL = rho_L * (S_L - u_L)
    dR = rho_R * (S_R - u_R)
    return ((p_R - p_L) + (u_L * dL - u_R * dR)) / (dL - dR)


def _hllc_star_flux(p, c, S_K, S_star, axis, left):
    """HLLC.py:41-78."""
    ua = 1 + axis
    m0, m1 = MINOR_AXES[axis]
    pre = (S_K - p[ua]) / (S_K - S_star) * p[0]
    us = [pre, pre, pre, pre,
          pre * (c[4] / c[0] + (S_star - p[ua]) * (S_star + p[4] / p[0] / (S_K - p[ua])))]
    us[ua] = us[ua] * S_star
    us[m0] = us[m0] * p[m0]
    us[m1] = us[m1] * p[m1]
    us = np.stack(us, axis=0)
    f = physical_flux(p, c, axis)
    S = np.minimum(S_K, 0.0) if left else np.maximum(S_K, 0.0)
    return f + S * (us - c)


def hllc(pl, pr, cl, cr, axis, gamma, signal_speed="EINFELDT"):
    """HLLC.py:80-126."""
    ua = 1 + axis
    aL = speed_of_sound(pl[4], pl[0], gamma)
    aR = speed_of_sound(pr[4], pr[0], gamma)
    S_L, S_R = signal_speeds(signal_speed, pl[ua], pr[ua], aL, aR, pl[0], pr[0], pl[4], pr[4], gamma)
    S_s = sstar(pl[ua], pr[ua], pl[4], pr[4], pl[0], pr[0], S_L, S_R)
    fL = _hllc_star_flux(pl, cl, S_L, S_s, axis, True)
    fR = _hllc_star_flux(pr, cr, S_R, S_s, axis, False)
    return 0.5 * (1 + np.sign(S_s)) * fL + 0.5 * (1 - np.sign(S_s)) * fR


def hll(pl, pr, cl, cr, axis, gamma, signal_speed="EINFELDT"):
    """solvers/riemann_solvers/HLL.py (single phase)."""
    ua = 1 + axis
    aL = speed_of_sound(pl[4], pl[0], gamma)
    aR = speed_of_sound(pr[4], pr[0], gamma)
    S_L, S_R = signal_speeds(signal_speed, pl[ua], pr[ua], aL, aR, pl[0], pr[0], pl[4], pr[4], gamma)
    wL = np.minimum(S_L, 0.0)
    wR = np.maximum(S_R, 0.0)
    fL = physical_flux(pl, cl, axis)
    fR = physical_flux(pr, cr, axis)
    return (wR * fL - wL * fR + wL * wR * (cr - cl)) / (wR - wL + EPS)


def hllclm(pl, pr, cl, cr, axis, gamma, signal_speed="EINFELDT"):
    """HLLCLM.py:30-135 (single phase): HLLC with the low-Mach wave-speed limiter of Fleischmann et al. 2020,
    Ma_limit = 0.1 (:28)."""
    ua = 1 + axis
    m0, m1 = MINOR_AXES[axis]
    aL = speed_of_sound(pl[4], pl[0], gamma)
    aR = speed_of_sound(pr[4], pr[0], gamma)
    S_L, S_R = signal_speeds(signal_speed, pl[ua], pr[ua], aL, aR, pl[0], pr[0], pl[4], pr[4], gamma)
    S_s = sstar(pl[ua], pr[ua], pl[4], pr[4], pl[0], pr[0], S_L, S_R)

    def ustar(p, c, S_K):                                                          # Toro 10.73, :96-109
        pre = (S_K - p[ua]) / (S_K - S_s) * p[0]
        us = [pre, pre, pre, pre,
              pre * (c[4] / c[0] + (S_s - p[ua]) * (S_s + p[4] / p[0] / (S_K - p[ua])))]
        us[ua] = us[ua] * S_s
        us[m0] = us[m0] * p[m0]
        us[m1] = us[m1] * p[m1]
        return np.stack(us, axis=0)
    usL, usR = ustar(pl, cl, S_L), ustar(pr, cr, S_R)
    Ma_local = np.maximum(np.abs(pl[ua] / aL), np.abs(pr[ua] / aR))                # :52-56
    phi = np.sin(np.minimum(1.0, Ma_local / 0.1) * np.pi * 0.5)
    wL = phi * S_L
    wR = phi * S_R
    fL = physical_flux(pl, cl, axis)
    fR = physical_flux(pr, cr, axis)
    flux_star = 0.5 * (fL + fR) + 0.5 * (wL * (usL - cl) + np.abs(S_s) * (usL - usR) + wR * (usR - cr))   # Eq. 19
    return (0.5 * (1 + np.sign(S_L)) * fL + 0.5 * (1 - np.sign(S_R)) * fR
            + 0.25 * (1 - np.sign(S_L)) * (1 + np.sign(S_R)) * flux_star)          # Eq. 18


def ausmp(pl, pr, cl, cr, axis, gamma):
    """AUSMP.py:29-95 (AUSM+, interface speed of sound ARITHMETIC, alpha = 3/16, beta = 1/8)."""
    ua = 1 + axis
    alpha, beta = 3.0 / 16.0, 1.0 / 8.0
    phi_L = np.stack([cl[0], cl[1], cl[2], cl[3], cl[4] + pl[4]], axis=0)          # get_phi :86-95
    phi_R = np.stack([cr[0], cr[1], cr[2], cr[3], cr[4] + pr[4]], axis=0)
    aL = speed_of_sound(pl[4], pl[0], gamma)
    aR = speed_of_sound(pr[4], pr[0], gamma)
    a = 0.5 * (aL + aR)
    M_l = pl[ua] / a
    M_r = pr[ua] / a
    M_plus = np.where(np.abs(M_l) >= 1, 0.5 * (M_l + np.abs(M_l)),
                      0.25 * np.square(M_l + 1.0) + beta * np.square(M_l * M_l - 1.0))
    M_minus = np.where(np.abs(M_r) >= 1, 0.5 * (M_r - np.abs(M_r)),
                       -0.25 * np.square(M_r - 1.0) - beta * np.square(M_r * M_r - 1.0))
    M_ausm = M_plus + M_minus
    M_ausm_plus = 0.5 * (M_ausm + np.abs(M_ausm))
    M_ausm_minus = 0.5 * (M_ausm - np.abs(M_ausm))
    P_plus = np.where(np.abs(M_l) >= 1.0, 0.5 * (1 + np.sign(M_l)),
                      0.25 * np.square(M_l + 1.0) * (2.0 - M_l) + alpha * M_l * np.square(M_l * M_l - 1.0))
    P_minus = np.where(np.abs(M_r) >= 1.0, 0.5 * (1 - np.sign(M_r)),
                       0.25 * np.square(M_r - 1.0) * (2.0 + M_r) - alpha * M_r * np.square(M_r * M_r - 1.0))
    pressure_ausm = P_plus * pl[4] + P_minus * pr[4]
    F = a * (M_ausm_plus * phi_L + M_ausm_minus * phi_R)
    F[ua] = F[ua] + pressure_ausm
    return F


def rusanov(pl, pr, cl, cr, axis, gamma):
    """Rusanov.py:25-47."""
    ua = 1 + axis
    aL = speed_of_sound(pl[4], pl[0], gamma)
    aR = speed_of_sound(pr[4], pr[0], gamma)
    alpha = np.maximum(np.abs(pl[ua]) + aL, np.abs(pr[ua]) + aR)
    return 0.5 * (physical_flux(pl, cl, axis) + physical_flux(pr, cr, axis)) - 0.5 * alpha * (cr - cl)


# --------------------------------------------------------------------------
# dissipative fluxes (CENTRAL4 stencils): solvers/source_term_solver.py
# --------------------------------------------------------------------------
def temperature(prims, s: Setup):
    """ideal_gas.py:64-65: T = p / (rho R), on whatever region is passed."""
    return prims[4] / (prims[0] * s.gas_constant)


def _sl(s: Setup, axis, lo, hi, ext):
    """[..., transverse = interior widened by `ext`, axis = lo:hi]."""
    nh = s.nh
    out = [Ellipsis]
    for i in range(3):
        if i == axis:
            out.append(slice(lo, hi if hi != 0 else None))
        elif s.cells[i] > 1:
            out.append(slice(nh - ext, -(nh - ext)))
        else:
            out.append(slice(None))
    return tuple(out)


def central4_reconstruct(buf, axis, s: Setup, n=None):
    """stencils/reconstruction/central/central_4.py:32-47 with spatial_stencil.py:99-111: faces f = 0..N of
    `axis`, cells n-2+f .. n+1+f, transverse range [n:-n] (n = the buffer's halo width)."""
    n = s.nh if n is None else n
    ext = s.nh - n            # only used with buffers that carry s.nh halos
    b = [buf[_sl(s, axis, n + k, -n + k + 1, ext)] for k in (-2, -1, 0, 1)]
    c0, c1 = -1.0 / 16.0, 9.0 / 16.0
    return c0 * (b[0] + b[3]) + c1 * (b[1] + b[2])


def deriv4_face(buf, dxi, axis, s: Setup):
    """stencils/derivative/deriv_face_4.py: 1/dx * (1/24 (u_{i-1} - u_{i+2}) + 27/24 (u_{i+1} - u_i))."""
    n = s.nh
    b = [buf[_sl(s, axis, n + k, -n + k + 1, 0)] for k in (-2, -1, 0, 1)]
    c0, c1 = 1.0 / 24.0, 27.0 / 24.0
    return 1.0 / dxi * (c0 * (b[0] - b[3]) + c1 * (b[2] - b[1]))


def deriv4_center_offset2(buf, dxi, axis, s: Setup):
    """stencils/derivative/deriv_center_4.py with nh = s.nh, offset = 2 (source_term_solver.py:63-70):
    cell-centre derivative on the interior widened by 2 cells in every active axis."""
    n = s.nh - 2
    b = [buf[_sl(s, axis, n + k, -n + k, 2)] for k in (-2, -1, 1, 2)]
    c0, c1 = 1.0 / 12.0, 8.0 / 12.0
    return 1.0 / dxi * (c0 * (b[0] - b[3]) + c1 * (b[2] - b[1]))


def _reconstruct_offset2(buf2, axis, s: Setup):
    """central_4 on a buffer that carries 2 halo cells (reconstruct_stencil_duidxi, source_term_solver.py:75-80)."""
    def sl(lo, hi):
        out = [Ellipsis]
        for i in range(3):
            if i == axis:
                out.append(slice(lo, hi if hi != 0 else None))
            elif s.cells[i] > 1:
                out.append(slice(2, -2))
            else:
                out.append(slice(None))
        return tuple(out)
    b = [buf2[sl(2 + k, -2 + k + 1)] for k in (-2, -1, 0, 1)]
    c0, c1 = -1.0 / 16.0, 9.0 / 16.0
    return c0 * (b[0] + b[3]) + c1 * (b[1] + b[2])


def _flux_shape(axis, s: Setup):
    return tuple(s.cells[i] + (1 if i == axis else 0) for i in range(3))


def _temperature_at_face(T, axis, s: Setup):
    T_cf = central4_reconstruct(T, axis, s)
    return np.where(T_cf <= 0.0, EPS, T_cf)                              # source_term_solver.py:289-292


def _dynamic_viscosity(T, s: Setup):
    """material.py:91-92 with the CUSTOM float wrapper (input/setup_reader.py:203-216): ones_like(T) * value,
    non-dimensionalised by rho_ref u_ref L_ref = 1."""
    return np.ones_like(T) * s.dynamic_viscosity


def _thermal_conductivity(T, s: Setup):
    if s.thermal_conductivity_model == "CUSTOM":
        return np.ones_like(T) * s.thermal_conductivity
    if s.thermal_conductivity_model == "PRANDTL":                        # material.py:115-116
        return s.cp * _dynamic_viscosity(T, s) / s.prandtl_number
    raise NotImplementedError(s.thermal_conductivity_model)


def viscous_flux_axis(prims, T, axis, s: Setup):
    """source_term_solver.py:258-345 (+ :405-470 velocity gradient, :503-533 tau, :535-582 derivatives):
    (4, faces) = (tau_axis0, tau_axis1, tau_axis2, u.tau)."""
    T_cf = _temperature_at_face(T, axis, s)
    mu_1 = _dynamic_viscosity(T_cf, s)
    mu_2 = s.bulk_viscosity - 2.0 / 3.0 * mu_1
    vel = prims[1:4]
    shape = _flux_shape(axis, s)
    grad = []                                  # grad[i] = d(vel)/dx_i at the faces of `axis`, (3, faces)
    for i in range(3):
        if i in s.active:
            if i == axis:
                g = deriv4_face(vel, s.dx[i], i, s)
            else:
                g = _reconstruct_offset2(deriv4_center_offset2(vel, s.dx[i], i, s), axis, s)
        else:
            g = np.zeros((3,) + shape)
        grad.append(g)
    vg = np.stack(grad, axis=1)                # vg[c, i] = d u_c / d x_i
    tau = []
    for k in range(3):
        if axis in s.active and k in s.active:
            tau.append(mu_1 * (vg[axis, k] + vg[k, axis]))
        else:
            tau.append(np.zeros(shape))
    tau[axis] = tau[axis] + mu_2 * sum([vg[k, k] for k in s.active])
    vel_cf = central4_reconstruct(vel, axis, s)
    if s.is_viscous_heat_production:
        vt = 0.0
        for k in s.active:
            vt = vt + tau[k] * vel_cf[k]
    else:
        vt = np.zeros_like(tau[0])
    return np.stack([*tau, vt])


def heat_flux_axis(prims, T, axis, s: Setup):
    """source_term_solver.py:188-250: q = -lambda(T_face) dT/dx_axis."""
    T_cf = _temperature_at_face(T, axis, s)
    lam = _thermal_conductivity(T_cf, s)
    return -lam * deriv4_face(T, s.dx[axis], axis, s)


# --------------------------------------------------------------------------
# right-hand side
# --------------------------------------------------------------------------
def flux_limiter(F, prims, cons, dt, axis, s: Setup):
    """limiter_flux.py:146-330 (SINGLE-PHASE, flux_limiter SIMPLE | NASA): a face whose high-order flux would drive
    the density (first check) or then the pressure (second check) of one of its two cells below eps under the
    pseudo-integration U -/+ 2 lambda F falls back to the first-order HLLC / Einfeldt flux (binary switch)."""
    eps_density, eps_pressure = 1e-12, 1e-10                            # config/precision.py:55
    nh, n = s.nh, s.cells[axis]

    def cells(buf, lo):                                                 # cons_positivity_slices, :134-138
        sl = [slice(None)] + list(s.interior)
        sl[1 + axis] = slice(lo, lo + n + 1)
        return buf[tuple(sl)]

    if s.flux_partition == "UNIFORM":                                   # compute_partition, :681-720
        one_sigma = len(s.active)
    elif s.flux_partition == "CELLSIZE":
        one_sigma = sum(s.inv_dx[a] for a in s.active) / s.inv_dx[axis]
    else:
        raise NotImplementedError(s.flux_partition)
    lam = dt * s.inv_dx[axis] * one_sigma                               # :205
    # first-order flux: WENO1 in PRIMITIVE variables + HLLC + Einfeldt (:58-99)
    pL, pR = cells(prims, nh - 1), cells(prims, nh)
    F_pos = hllc(pL, pR, cons_from_prims(pL, s.gamma), cons_from_prims(pR, s.gamma), axis, s.gamma, "EINFELDT")
    c_plus, c_minus = cells(cons, nh - 1), cells(cons, nh)              # cell i (+), cell i+1 (-)
    if s.flux_limiter == "NASA":
        fs = physical_flux(prims, cons, axis)                           # equation_manager.get_fluxes_xi
        fs_plus, fs_minus = cells(fs, nh - 1), cells(fs, nh)

    def integrate(F):                                                   # integrate_positivity, :471-560
        if s.flux_limiter == "SIMPLE":
            return c_minus + 2.0 * lam * F, c_plus - 2.0 * lam * F
        if s.flux_limiter == "NASA":
            return c_minus + 2.0 * lam * (F - fs_minus), c_plus - 2.0 * lam * (F - fs_plus)
        raise NotImplementedError(s.flux_limiter)

    U_minus, U_plus = integrate(F)
    theta = np.where(np.minimum(U_minus[0], U_plus[0]) < eps_density, 1, 0)            # :571-575
    F = theta * F_pos + (1 - theta) * F                                                # :766
    U_minus, U_plus = integrate(F)
    W_minus, W_plus = prims_from_cons(U_minus, s.gamma), prims_from_cons(U_plus, s.gamma)
    theta = np.where(np.minimum(W_minus[4] + 0.0, W_plus[4] + 0.0) < eps_pressure, 1, 0)   # :620-628 (pb = 0)
    return theta * F_pos + (1 - theta) * F


def frozen_state(pL, pR, gamma, kind="ARITHMETIC"):
    """eigendecomposition.py:120-281 (single phase, ideal gas): (primes_ave, enthalpy_ave, grueneisen_ave, c_ave, cc_ave,
    velocity_square) at the face between the cells with primitives pL, pR."""
    def total_enthalpy(p):                                                     # ideal_gas.py:90-110
        E = p[4] / (gamma - 1) + 0.5 * p[0] * (np.square(p[1]) + np.square(p[2]) + np.square(p[3]))
        return (E + p[4]) / p[0]
    if kind == "ARITHMETIC":                                                   # :146-231
        ave = 0.5 * (pL + pR)
        G = (gamma - 1) * np.ones_like(ave[0])                                 # get_grueneisen, ideal_gas.py:65-67
        H = total_enthalpy(ave)
        c = np.sqrt(gamma * ave[4] / ave[0])
        cc = c * c
        q2 = np.sum(np.square(ave[1:4]), axis=0)
    elif kind == "ROE":                                                        # :233-276, compute_roe_cons :283-294
        ave = (np.sqrt(pL[0]) * pL + np.sqrt(pR[0]) * pR) / (np.sqrt(pL[0]) + np.sqrt(pR[0]))
        ave[0] = np.sqrt(pL[0] * pR[0])
        sL, sR = np.sqrt(pL[0]), np.sqrt(pR[0])
        rho_div = 1.0 / (sL + sR)
        H = (sL * total_enthalpy(pL) + sR * total_enthalpy(pR)) * rho_div
        psi = (sL * (pL[4] / pL[0]) + sR * (pR[4] / pR[0])) * rho_div          # get_psi, ideal_gas.py:61-63
        G = (sL * (gamma - 1) + sR * (gamma - 1)) * rho_div
        dq2 = np.sum(np.square(pR[1:4] - pL[1:4]), axis=0)
        p_over_rho = (sL * pL[4] / pL[0] + sR * pR[4] / pR[0]) * rho_div + 0.5 * ave[0] * rho_div * rho_div * dq2
        q2 = np.sum(np.square(ave[1:4]), axis=0)
        cc = psi + G * p_over_rho
        c = np.sqrt(cc)
    else:
        raise NotImplementedError(kind)
    return ave, H, G, c, cc, q2


def _matvec(M, x):
    """jnp.einsum("ij...,j...->i...", M, x) (eigendecomposition.py:717-743) with the sum over j taken in order."""
    out = []
    for i in range(5):
        acc = M[i][0] * x[0]
        for j in range(1, 5):
            acc = acc + M[i][j] * x[j]
        out.append(acc)
    return np.stack(out, axis=0)


def conservative_eigensystem(pL, pR, axis, gamma, flux_splitting=None, frozen="ARITHMETIC"):
    """eigendecomposition.py:576-715 at the frozen state of :120-281 (single phase, ideal gas): right and
    left eigenvectors of the conservative flux Jacobian after Fedkiw et al. 1999 as 5x5 nested lists of face arrays,
    and -- for the flux-splitting scheme -- the eigenvalue magnitudes (ROE :668-671, CLLF :674-681, LLF :684-689)."""
    ua = 1 + axis
    m0, m1 = MINOR_AXES[axis]
    ave, H, G, c, cc, q2 = frozen_state(pL, pR, gamma, frozen)
    one_cc = 1.0 / cc
    one_rho = 1.0 / ave[0]
    Z = np.zeros_like(ave[0])
    Rm = [[Z for _ in range(5)] for _ in range(5)]
    Lm = [[Z for _ in range(5)] for _ in range(5)]
    Rm[0][0] = np.ones_like(Z)
    Rm[ua][0] = ave[ua] - c
    Rm[m0][0] = ave[m0]
    Rm[m1][0] = ave[m1]
    Rm[4][0] = H - ave[ua] * c
    Rm[0][ua] = G
    Rm[1][ua] = G * ave[1]
    Rm[2][ua] = G * ave[2]
    Rm[3][ua] = G * ave[3]
    Rm[4][ua] = G * H - cc
    Rm[m0][m0] = -ave[0]
    Rm[4][m0] = -ave[0] * ave[m0]
    Rm[m1][m1] = ave[0]
    Rm[4][m1] = ave[0] * ave[m1]
    Rm[0][4] = np.ones_like(Z)
    Rm[ua][4] = ave[ua] + c
    Rm[m0][4] = ave[m0]
    Rm[m1][4] = ave[m1]
    Rm[4][4] = H + ave[ua] * c
    Lm[0][0] = 0.5 * one_cc * (G * q2 - G * H + (ave[ua] + c) * c)
    Lm[0][ua] = 0.5 * one_cc * (-ave[ua] * G - c)
    Lm[0][m0] = 0.5 * one_cc * (-ave[m0] * G)
    Lm[0][m1] = 0.5 * one_cc * (-ave[m1] * G)
    Lm[0][4] = 0.5 * one_cc * G
    Lm[ua][0] = one_cc * (H - q2)
    Lm[ua][1] = ave[1] * one_cc
    Lm[ua][2] = ave[2] * one_cc
    Lm[ua][3] = ave[3] * one_cc
    Lm[ua][4] = -one_cc
    Lm[m0][0] = ave[m0] * one_rho
    Lm[m0][m0] = -one_rho
    Lm[m1][0] = -ave[m1] * one_rho
    Lm[m1][m1] = one_rho
    Lm[4][0] = 0.5 * one_cc * (G * q2 - G * H - (ave[ua] - c) * c)
    Lm[4][ua] = 0.5 * one_cc * (-ave[ua] * G + c)
    Lm[4][m0] = 0.5 * one_cc * (-ave[m0] * G)
    Lm[4][m1] = 0.5 * one_cc * (-ave[m1] * G)
    Lm[4][4] = 0.5 * one_cc * G
    lam = None
    if flux_splitting == "ROE":
        lam = (np.abs(ave[ua] - c), np.abs(ave[ua]), np.abs(ave[ua] + c))
    elif flux_splitting == "CLLF":
        cL, cR = speed_of_sound(pL[4], pL[0], gamma), speed_of_sound(pR[4], pR[0], gamma)
        lam = (np.maximum(np.abs(pL[ua] - cL), np.abs(pR[ua] - cR)), np.maximum(np.abs(pL[ua]), np.abs(pR[ua])),
               np.maximum(np.abs(pL[ua] + cL), np.abs(pR[ua] + cR)))
    elif flux_splitting == "LLF":
        g = np.maximum(np.abs(pL[ua]) + speed_of_sound(pL[4], pL[0], gamma),
                       np.abs(pR[ua]) + speed_of_sound(pR[4], pR[0], gamma))
        lam = (g, g, g)
    elif flux_splitting is not None:
        raise NotImplementedError(flux_splitting)
    return Rm, Lm, lam


def flux_splitting_face_flux(prims, cons, axis, s: Setup):
    """flux_splitting_scheme.py:62-111: conservatives and physical fluxes of the stencil cells in the characteristic space
    of the face's frozen state, split by the eigenvalue magnitudes, F+ reconstructed from the left (j = 0), F- from
    the right (j = 1), summed and transformed back."""
    wp = _window(prims, axis, s)
    wc = _window(cons, axis, s)
    Rm, Lm, lam = conservative_eigensystem(wp[2], wp[3], axis, s.gamma, s.flux_splitting, s.frozen_state)
    lam5 = (lam[0], lam[1], lam[1], lam[1], lam[2])
    pos, neg = [], []
    for p, c in zip(wp, wc):
        char = _matvec(Lm, c)
        char_flux = _matvec(Lm, physical_flux(p, c, axis))
        lam_char = np.stack([lam5[i] * char[i] for i in range(5)], axis=0)
        pos.append(0.5 * (char_flux + lam_char))
        neg.append(0.5 * (char_flux - lam_char))
    stencil = STENCILS[s.stencil]
    char_flux_xi = stencil(pos, 0) + stencil(neg[::-1], 1)
    return _matvec(Rm, char_flux_xi)


def face_flux(prims, axis, s: Setup, cons=None, dt=None):
    """high_order_godunov.py:117-231: face fluxes (5, N_axis+1, transverse interior); with positivity/flux_limiter
    the positivity-preserving switch of space_solver.py:532-543 follows."""
    if s.convective_solver == "FLUX-SPLITTING":
        cons = cons_from_prims(prims, s.gamma) if cons is None else cons
        return flux_splitting_face_flux(prims, cons, axis, s)
    pl, pr, cl, cr = reconstruct(prims, axis, s, cons)
    if s.riemann == "HLLC":
        F = hllc(pl, pr, cl, cr, axis, s.gamma, s.signal_speed)
    elif s.riemann == "RUSANOV":
        F = rusanov(pl, pr, cl, cr, axis, s.gamma)
    elif s.riemann == "HLL":
        F = hll(pl, pr, cl, cr, axis, s.gamma, s.signal_speed)
    elif s.riemann == "HLLC-LM":
        F = hllclm(pl, pr, cl, cr, axis, s.gamma, s.signal_speed)
    elif s.riemann == "AUSMP":
        F = ausmp(pl, pr, cl, cr, axis, s.gamma)
    else:
        raise NotImplementedError(s.riemann)
    if s.flux_limiter:
        if dt is None:
            raise ValueError("positivity/flux_limiter needs the physical time step size")
        cons = cons_from_prims(prims, s.gamma) if cons is None else cons
        with np.errstate(all="ignore"):
            F = flux_limiter(F, prims, cons, dt, axis, s)
    return F


def rhs_axis(prims, axis, s: Setup, cons=None, dt=None):
    """space_solver.py:456-674 (convective branch: :489, :517-543, :597-599)."""
    if s.is_convective_flux:
        fc = face_flux(prims, axis, s, cons, dt)
        f = np.zeros_like(fc) + fc                                      # :517, :545
    else:
        f = np.zeros((5,) + _flux_shape(axis, s))                       # :517 only
    if s.is_dissipative:
        T = temperature(prims, s)
        if s.is_viscous_flux:                                           # :567-573
            f = f.copy()
            f[1:5] = f[1:5] + (-viscous_flux_axis(prims, T, axis, s))
        if s.is_heat_flux:                                              # :576-584
            f = f.copy()
            f[4] = f[4] + heat_flux_axis(prims, T, axis, s)
    lo = [slice(None)] * 4
    hi = [slice(None)] * 4
    lo[1 + axis] = slice(None, -1)
    hi[1 + axis] = slice(1, None)
    out = np.zeros((5,) + s.cells)
    out = out + s.inv_dx[axis] * (f[tuple(lo)] - f[tuple(hi)])
    return out


def gravity_forces(cons, s: Setup):
    """source_term_solver.py:163-186: (4, ...) = (g_i rho, g . (rho u)) on the interior, with the reference's einsums."""
    ci = cons[(slice(None),) + s.interior]
    g = np.array(s.gravity, dtype=np.float64)
    density = np.expand_dims(ci[0], axis=0)
    momentum = ci[1:4]
    mom = np.einsum("ij..., jk...->ik...", g.reshape(3, 1), density)
    ene = np.einsum("ij..., jk...->ik...", g.reshape(1, 3), momentum)
    return np.concatenate([mom, ene], axis=0)


def compute_rhs(prims, s: Setup, cons=None, dt=None):
    """space_solver.py:151-453 (single phase): 0.0 + rhs_x + rhs_y + rhs_z (+ volume forces, :378-384, which
    read the conservatives)."""
    rhs = 0.0
    for axis in s.active:
        rhs = rhs + rhs_axis(prims, axis, s, cons, dt)
    if s.is_volume_force:
        cons = cons_from_prims(prims, s.gamma) if cons is None else cons
        rhs = rhs.copy()
        rhs[1:5] = rhs[1:5] + gravity_forces(cons, s)
    return rhs


# --------------------------------------------------------------------------
# time step size, positivity info
# --------------------------------------------------------------------------
def time_step_size(prims, s: Setup):
    """time_integration/time_step_size.py:15-157 (convective contribution only)."""
    if s.fixed_timestep:
        return float(s.fixed_timestep)
    pi = prims[(slice(None),) + s.interior]
    c = speed_of_sound(pi[4], pi[0], s.gamma)
    acc = 0.0
    for i in s.active:
        acc = acc + (np.abs(pi[1 + i]) + c)
    dt = s.dx_min / (np.max(acc) + EPS)
    if s.is_dissipative:
        T = temperature(pi, s)
        dx2 = s.dx_min * s.dx_min
        if s.is_viscous_flux:                                           # time_step_size.py:111-121
            nu = _dynamic_viscosity(T, s) / pi[0]
            dt = np.minimum(dt, 3.0 / 14.0 * dx2 / (np.max(nu) + EPS))
        if s.is_heat_flux:                                              # :123-135
            alpha = _thermal_conductivity(T, s) / (pi[0] * s.cp)
            dt = np.minimum(dt, 0.1 * dx2 / (np.max(alpha) + EPS))
    dt = dt * s.cfl
    return float(dt)


def positivity_info(prims, s: Setup):
    """solvers/positivity/positivity_handler.py:245-254: (min rho, min p) over the interior."""
    pi = prims[(slice(None),) + s.interior]
    return float(np.min(pi[0])), float(np.min(pi[4]))


# --------------------------------------------------------------------------
# initialisation and the step
# --------------------------------------------------------------------------
def initialize(prims_interior, s: Setup, from_user_buffer=False):
    """initialization/material_fields_initializer.py:590-690 (IC evaluated on the mesh)
    or :148-210/:425-497 (user array).  `prims_interior` is (5,Nx,Ny,Nz); with
    from_user_buffer=True it is (5-3+dim, ...) and inactive velocities keep the
    eps fill of the buffer (helper_functions.py:21-60), as in the reference."""
    buf = np.ones(s.shape) * EPS
    sl = (slice(None),) + s.interior
    if from_user_buffer:
        idx = [0] + [1 + i for i in s.active] + [4]
        buf[(idx,) + s.interior] = prims_interior
    else:
        buf[sl] = prims_interior
    cons = cons_from_prims(buf, s.gamma)
    return halo_fill(buf, cons, s)


def stage(prims, cons, cons_n, dt, k, s: Setup):
    """One RK stage: simulation_manager.py:770-1047 (single-phase branch).
    Returns (prims, cons, rhs)."""
    rk = RK[s.integrator]
    rhs = compute_rhs(prims, s, cons, dt)                               # :796 (dt: the flux limiter's lambda)
    if k > 0:                                                           # RK3.py:49-50
        a, b = rk["blend"][k - 1]
        cons = a * cons + b * cons_n
    step = dt * rk["dt_mult"][k]                                        # RK3.py:60
    cons = cons.copy()
    sl = (slice(None),) + s.interior
    cons[sl] = cons[sl] + step * rhs                                    # time_integrator.py:57
    prims = prims_from_cons(cons, s.gamma)                              # :943
    prims, cons = halo_fill(prims, cons, s)                             # :963
    return prims, cons, rhs


def step(prims, cons, dt, s: Setup, record=None):
    """One full time step; returns (prims, cons, dt_next)."""
    cons_n = cons
    with np.errstate(all="ignore"):     # untouched corner cells hold eps/garbage, as in the reference
        for k in range(RK[s.integrator]["stages"]):
            prims, cons, rhs = stage(prims, cons, cons_n, dt, k, s)
            if record is not None:
                record["rhs"].append(rhs)
                record["prims"].append(prims)
                record["cons"].append(cons)
    return prims, cons, time_step_size(prims, s)                        # simulation_manager.py:612-628


def totals(cons, s: Setup):
    ci = cons[(slice(None),) + s.interior]
    return np.array([ci[v].sum() for v in range(5)])
