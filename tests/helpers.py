"""Shared test helpers: setups, fixtures, error metrics."""
import copy
import glob
import json
import os

import numpy as np

from oracle import port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# north_star tolerances
TOL_RHS = 1e-12        # per-stage RHS, relative L-inf
TOL_PRIMS_100 = 1e-9   # primitives after 100 steps, relative L-inf
TOL_TOTALS = 1e-12     # conserved totals


def golden_names(dissipative=None):
    """All fixtures; dissipative=False/True keeps only the convective-only / the viscous+heat ones."""
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
    if dissipative is None:
        return names
    # "dissipative" = anything beyond the window -> face flux map in the rhs (viscous / heat flux, gravity, and the
    # flux limiter, whose rhs also depends on the time step)
    return [n for n in names if ("visc" in n or "gravity" in n or "noconv" in n or "fluxlim" in n) == bool(dissipative)]


def generic_golden_names():
    """Fixtures of the generic reconstruction stencils / RK2_LS4 (tests/golden/generic/), as names load_golden takes."""
    return sorted("generic/" + os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "generic", "*.npz")))


def api_golden_names():
    """Fixtures that only the public API can run (tests/golden/api/): setups whose data lives in the host runtime."""
    return sorted("api/" + os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "api", "*.npz")))


GENERIC_STENCILS = ("WENO1", "WENO3-JS", "WENO3-Z", "WENO3-N", "CENTRAL2", "TENO5", "TENO5-A", "TENO6", "TENO6-A", "WENO6-CU",
                    "KOREN", "MC", "MINMOD", "SUPERBEE", "VANALBADA", "VANLEER")


def load_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    case = json.loads(str(g["case_json"]))
    num = json.loads(str(g["num_json"]))
    return g, case, num


def dissipation_from_json(case, num) -> dict:
    """active_physics + material_properties/transport -> the oracle Setup's dissipative-flux fields."""
    ap = num.get("active_physics", {})
    if not (ap.get("is_viscous_flux") or ap.get("is_heat_flux")):
        return {}
    tr = case["material_properties"].get("transport", {}) or {}
    mu = tr.get("dynamic_viscosity", {}) or {}
    tc = tr.get("thermal_conductivity", {}) or {}
    return dict(is_viscous_flux=bool(ap.get("is_viscous_flux")), is_heat_flux=bool(ap.get("is_heat_flux")),
                is_viscous_heat_production=bool(ap.get("is_viscous_heat_production", True)),
                dynamic_viscosity=float(mu.get("value", 0.0)), bulk_viscosity=float(tr.get("bulk_viscosity", 0.0)),
                thermal_conductivity_model=tc.get("model", "CUSTOM"),
                thermal_conductivity=float(tc.get("value", 0.0) or 0.0),
                prandtl_number=float(tc.get("prandtl_number", 1.0) or 1.0),
                gas_constant=float(case["material_properties"]["equation_of_state"]["specific_gas_constant"]))


def _transverse_mesh(case, face):
    d = case["domain"]
    cells = [d[a]["cells"] for a in "xyz"]
    ax = port.FACE_AXIS[face]
    trans = [i for i in range(3) if i != ax and cells[i] > 1]
    centers = []
    for i in trans:
        lo, hi = d["xyz"[i]]["range"]
        dx = (hi - lo) / cells[i]
        centers.append(np.linspace(lo + dx / 2, hi - dx / 2, cells[i]))
    mesh = np.meshgrid(*centers, indexing="ij") if centers else []
    return mesh, [cells[i] if i in trans else 1 for i in range(3)]


def multi_type_entries(case, face):
    """A face given as a list of types with bounding_domain lambdas (halos/outer/material.py:121-277)."""
    mesh, shape = _transverse_mesh(case, face)
    out = []
    for ent in case["boundary_conditions"][face]:
        mask = np.asarray(eval(ent["bounding_domain"], {"jnp": np, "np": np})(*mesh)).reshape(shape)   # noqa: S307
        vals = dirichlet_values(case, face, entry=ent) if ent["type"] == "DIRICHLET" else None
        out.append(dict(kind=ent["type"], mask=mask, values=vals))
    return out


def kernel_type_of(entry):
    """bc[face] of a multi-type face: the entry the kernels fill (the first that is not DIRICHLET)."""
    if isinstance(entry, list):
        return next(e["type"] for e in entry if e["type"] != "DIRICHLET")
    return entry["type"]


def dirichlet_values(case, face, keys=("rho", "u", "v", "w", "p"), entry=None, name="primitives_callable"):
    """primitives_callable of a DIRICHLET face (halos/outer/material.py:770-790, boundary_condition.py:105-126): floats, or
    lambdas of the ACTIVE transverse coordinates and the time, evaluated on the mesh grid of the face's transverse cell
    centres (single block) and shaped like the halo slab with extent 1 along the normal and the inactive axes."""
    d = case["domain"]
    cells = [d[a]["cells"] for a in "xyz"]
    ax = port.FACE_AXIS[face]
    trans = [i for i in range(3) if i != ax and cells[i] > 1]
    centers = []
    for i in trans:
        lo, hi = d["xyz"[i]]["range"]
        dx = (hi - lo) / cells[i]
        centers.append(np.linspace(lo + dx / 2, hi - dx / 2, cells[i]))
    mesh = np.meshgrid(*centers, indexing="ij") if centers else []
    shape = [cells[i] if i in trans else 1 for i in range(3)]
    out = []
    for k in ("rho", "u", "v", "w", "p"):
        if k not in keys:                     # SIMPLE_INFLOW takes no p, SIMPLE_OUTFLOW only p
            out.append(None)
            continue
        v = (entry or case["boundary_conditions"][face])[name][k]
        if isinstance(v, str):
            fn = eval(v, {"jnp": np, "np": np})                       # noqa: S307 -- the reference's own contract
            v = np.asarray(fn(*mesh, 0.0), dtype=np.float64).reshape(shape)
        else:
            v = float(v)
        out.append(v)
    return tuple(out)


def _type(case, face):
    """Type of a single-type face (None for a face given as a list of types)."""
    e = case["boundary_conditions"][face]
    return None if isinstance(e, list) else e["type"]


# entries of primitives_callable the type reads (read_boundary_conditions.py:160-365)
BC_VALUE_KEYS = {"NEUMANN": ("rho", "u", "v", "w", "p"), "SIMPLE_INFLOW": ("rho", "u", "v", "w"), "SIMPLE_OUTFLOW": ("p",)}


def setup_from_json(case, num) -> port.Setup:
    d = case["domain"]
    c = num["conservatives"]
    solver = c["convective_fluxes"].get("convective_solver", "GODUNOV")
    fs = c["convective_fluxes"].get("flux_splitting", {}) or {}
    g = c["convective_fluxes"].get("godunov", {}) or {}
    if solver == "FLUX-SPLITTING":          # the stencil of the flux_splitting block; the godunov block is not read
        g = dict(g, reconstruction_stencil=fs.get("reconstruction_stencil", "WENO5-Z"))
    return port.Setup(
        convective_solver=solver, flux_splitting=fs.get("flux_splitting", "ROE"),
        frozen_state=(fs if solver == "FLUX-SPLITTING" else g).get("frozen_state", "ARITHMETIC"),
        cells=tuple(d[a]["cells"] for a in "xyz"),
        domain=tuple(tuple(d[a]["range"]) for a in "xyz"),
        bc={f: kernel_type_of(case["boundary_conditions"][f]) for f in port.FACES},
        bc_multi={f: multi_type_entries(case, f) for f in port.FACES if isinstance(case["boundary_conditions"][f], list)},
        gamma=case["material_properties"]["equation_of_state"]["specific_heat_ratio"],
        nh=c["halo_cells"],
        recon=g.get("reconstruction_variable", "PRIMITIVE"),
        stencil=g.get("reconstruction_stencil", "WENO5-Z"),
        riemann=g.get("riemann_solver", "HLLC"),
        signal_speed=g.get("signal_speed", "EINFELDT"),
        integrator=c["time_integration"]["integrator"],
        cfl=c["time_integration"].get("CFL", 0.5),
        is_interpolation_limiter=bool((c.get("positivity", {}) or {}).get("is_interpolation_limiter", False)),
        limit_velocity=bool((c.get("positivity", {}) or {}).get("limit_velocity", False)),
        flux_limiter=(c.get("positivity", {}) or {}).get("flux_limiter", None) or None,
        flux_partition=(c.get("positivity", {}) or {}).get("flux_partition", "UNIFORM"),
        wall_velocity={f: dirichlet_values(case, f, ("u", "v", "w"), name="wall_velocity_callable")[1:4]
                       for f in port.FACES if _type(case, f) == "WALL"},
        dirichlet={f: dirichlet_values(case, f) for f in port.FACES if _type(case, f) == "DIRICHLET"},
        bc_values={f: dirichlet_values(case, f, BC_VALUE_KEYS[_type(case, f)]) for f in port.FACES
                   if _type(case, f) in BC_VALUE_KEYS},
        is_volume_force=bool(num.get("active_physics", {}).get("is_volume_force", False)),
        is_convective_flux=bool(num.get("active_physics", {}).get("is_convective_flux", True)),
        gravity=tuple(float(x) for x in (case.get("forcings", {}) or {}).get("gravity", (0.0, 0.0, 0.0))),
        **dissipation_from_json(case, num),
    )


def make_setup(cells, bc="PERIODIC", recon="CHAR-PRIMITIVE", riemann="HLLC", integrator="RK3", gamma=1.4,
               length=1.0, nh=5, stencil="WENO5-Z"):
    cells = tuple(cells)
    bcs = {}
    for f in port.FACES:
        ax = port.FACE_AXIS[f]
        bcs[f] = (bc if isinstance(bc, str) else bc[f]) if cells[ax] > 1 else "INACTIVE"
    return port.Setup(cells=cells, domain=((0.0, length),) * 3, bc=bcs, gamma=gamma, nh=nh, recon=recon,
                      riemann=riemann, integrator=integrator, stencil=stencil)


def smooth_ic(s: port.Setup, seed=0, amp=0.2):
    """Deterministic smooth-but-generic periodic primitives (5, Nx, Ny, Nz), plus a steep feature."""
    rng = np.random.default_rng(seed)
    x, y, z = np.meshgrid(*s.cell_centers(), indexing="ij")
    L = [s.domain[i][1] - s.domain[i][0] for i in range(3)]
    ph = rng.uniform(0, 2 * np.pi, size=(5, 3))
    k = 2 * np.pi

    def wave(v):
        out = 0.0
        for i, (c, n) in enumerate(zip((x, y, z), s.cells)):
            if n > 1:
                out = out + np.sin(k * c / L[i] + ph[v, i]) * np.cos(2 * k * c / L[i] - ph[v, (i + 1) % 3])
        return out
    rho = 1.0 + amp * wave(0)
    u = amp * 2 * wave(1)
    v = amp * 2 * wave(2) if s.cells[1] > 1 else np.zeros_like(rho)
    w = amp * 2 * wave(3) if s.cells[2] > 1 else np.zeros_like(rho)
    p = 1.0 + amp * wave(4)
    # a steep (but resolved-by-WENO) bump so the nonlinear weights are exercised
    r2 = sum(((c - 0.5 * (s.domain[i][0] + s.domain[i][1])) / L[i]) ** 2
             for i, (c, n) in enumerate(zip((x, y, z), s.cells)) if n > 1)
    rho = rho + 0.5 * (r2 < 0.04)
    p = p + 0.7 * (r2 < 0.04)
    return np.stack([rho, u, v, w, p]).astype(np.float64)


def field_scales(b, floor=1e-30):
    """Per-field normalisation: max|b_v|, floored at 1e-3 of the largest field (fields that are
    identically ~0 in the reference, e.g. w-momentum in TGV, cannot be normalised by themselves;
    SURVEY 8c)."""
    b = np.asarray(b)
    scale_all = max(float(np.max(np.abs(b))), floor)
    return np.array([max(float(np.max(np.abs(b[v]))), 1e-3 * scale_all) for v in range(b.shape[0])])


def rhs_scales(prims, s):
    """Per-field scale of a stage RHS: max over cells of sum_axes |rhs_axis| -- the size of the terms
    the RHS is a sum of.  At low Mach the total is a cancellation of large axis contributions (TGV:
    energy 169 + (-169) -> 0.3; mass 0.94 - 0.94 -> 1e-4), so normalising the error by the cancelled
    total would measure the conditioning of the reference's own formula (the reference evaluated with
    and without FMA contraction already differs by 1e-11 in that norm, tests/test_hostsim.py), not the
    kernel.  The north-star bound 1e-12 is applied in this norm."""
    if s.flux_limiter:           # the scale is a size, not a comparison: take it from the unlimited fluxes (no dt needed)
        s = copy.copy(s)
        s.flux_limiter = None
    tot = 0.0
    for a in s.active:
        tot = tot + np.abs(port.rhs_axis(prims, a, s))
    sc = np.asarray(field_scales(tot), dtype=float)
    # Floor: 1 % of the largest term the numerical flux is built from, per field: (1/dx) max(|F_v|, (|u_n| + c) |U_v|)
    # (HLLC = physical flux + S_K (U*_K - U_K)).  Each axis contribution is itself a difference of two face fluxes;
    # for a fluid (nearly) at rest with uniform pressure (lid-driven cavity at t = 0) those are equal O(p/dx),
    # O(rho c/dx) numbers, so the rounding floor of the rhs is eps * rho c/dx however small the rhs itself is.
    fl = np.zeros(5)
    pi = np.nan_to_num(prims[(slice(None),) + s.interior], nan=1.0)
    ci = np.abs(port.cons_from_prims(pi, s.gamma))
    c = port.speed_of_sound(pi[4], pi[0], s.gamma)
    for a in s.active:
        if not s.is_convective_flux:
            continue
        f = np.abs(np.nan_to_num(port.face_flux(prims, a, s)))
        wave = (np.abs(pi[1 + a]) + c)[None] * ci
        fl = np.maximum(fl, np.maximum(f.reshape(5, -1).max(axis=1), wave.reshape(5, -1).max(axis=1)) * float(s.inv_dx[a]))
    return np.maximum(sc, 1e-2 * fl)


def rel_linf(a, b, scale=None):
    """max_v max|a_v-b_v| / scale_v with scale_v = field_scales(b) unless given.  For a single
    axis' contribution to the RHS pass the scales of the TOTAL stage RHS: the north-star bound is on
    the stage RHS, and one axis can be ~0 (TGV: w = 0) while its fluxes are O(100)."""
    a, b = np.asarray(a), np.asarray(b)
    sc = field_scales(b) if scale is None else np.asarray(scale)
    return max(float(np.max(np.abs(a[v] - b[v]))) / float(sc[v]) for v in range(b.shape[0]))


def face_halo_mask(s: port.Setup):
    """Boolean mask (X,Y,Z) of cells the path defines: interior + face halos (no edges/corners)."""
    shape = s.shape[1:]
    inter = [np.zeros(n, bool) for n in shape]
    for i in range(3):
        if s.cells[i] > 1:
            inter[i][s.nh:-s.nh] = True
        else:
            inter[i][:] = True
    ix, iy, iz = np.meshgrid(*inter, indexing="ij")
    n_out = (~ix).astype(int) + (~iy).astype(int) + (~iz).astype(int)
    return n_out <= 1


def defined_mask(s: port.Setup):
    """Cells the path defines: interior + face halos, + edge halos with the viscous / heat flux."""
    shape = s.shape[1:]
    inter = [np.zeros(n, bool) for n in shape]
    for i in range(3):
        if s.cells[i] > 1:
            inter[i][s.nh:-s.nh] = True
        else:
            inter[i][:] = True
    ix, iy, iz = np.meshgrid(*inter, indexing="ij")
    n_out = (~ix).astype(int) + (~iy).astype(int) + (~iz).astype(int)
    return n_out <= (2 if s.is_dissipative else 1)


def apply_face_data_numpy(face_data, prims, cons, cells, nh, gamma):
    """What the halo kernel / the fused halo images do with jxf_set_face_data's arrays (sweep_kernels.cuh
    apply_face_data), on NumPy arrays that hold the faces' BASE-rule halos: per variable op 1 = replace by the data,
    op 2 = add it, inside the optional mask, the same for every halo layer; then the conservatives of those halo
    cells (equation_manager.py:93-101).  face_data: {face: (ops, data (5, n1, n2), mask (n1, n2) or None)} (tensors or
    arrays).  Used by the CPU tests of BlockRuntime's boundary-data construction."""
    prims, cons = prims.copy(), cons.copy()
    g1 = gamma - 1.0
    for face, (ops, data, mask) in face_data.items():
        data = np.asarray(data)
        mask = None if mask is None else np.asarray(mask).astype(bool)
        ax = port.FACE_AXIS[face]
        hi = face in ("east", "north", "top")
        idx = [slice(None)] + [slice(nh, -nh) if n > 1 else slice(None) for n in cells]
        idx[1 + ax] = slice(-nh, None) if hi else slice(0, nh)
        h = prims[tuple(idx)]                               # view (5, ..nh along ax..)
        shape = [n if n > 1 else 1 for n in cells]
        shape[ax] = 1
        for v in range(5):
            op = (ops >> (2 * v)) & 3
            if op == 0:
                continue
            d = data[v].reshape(shape)
            new = np.broadcast_to(d, h[v].shape) if op == 1 else h[v] + d
            h[v] = new if mask is None else np.where(mask.reshape(shape), new, h[v])
        with np.errstate(all="ignore"):
            e = h[4] / (h[0] * g1)
            c = cons[tuple(idx)]
            c[0] = h[0]
            c[1] = h[0] * h[1]
            c[2] = h[0] * h[2]
            c[3] = h[0] * h[3]
            c[4] = h[0] * (0.5 * (np.square(h[1]) + np.square(h[2]) + np.square(h[3])) + e)
    return prims, cons
