"""CPU, build container only: oracle/port.py is bit-identical to the reference's own sources run on
the NumPy `jax` stand-in.  Skipped where /root/reference does not exist (the GPU box)."""
import numpy as np
import pytest

from oracle.refharness import run_reference as rr

pytestmark = pytest.mark.skipif(not rr.reference_available(), reason="reference sources not present")


@pytest.mark.parametrize("name,kw,nsteps", [
    ("sod", dict(cells=(120, None, None)), 3),
    ("sod", dict(cells=(64, None, None), recon="PRIMITIVE", riemann="RUSANOV", integrator="EULER"), 2),
    ("riemann2d", dict(cells=(20, 28, None)), 2),
    ("tgv", dict(cells=(12, 12, 12)), 1),
    ("tgv", dict(cells=(10, 12, 14), bc="PERIODIC", recon="PRIMITIVE"), 1),
    # viscous + heat flux, edge halos (SYMMETRY mirror / PERIODIC copy / ANY_ANY mean)
    ("tgv", dict(cells=(10, 10, 12), dissipation=dict(mu=1 / 160, prandtl=0.71)), 1),
    ("tgv", dict(cells=(8, 10, 12), bc="PERIODIC", dissipation=dict(mu=0.01, bulk=0.002, kappa=0.05)), 1),
    ("riemann2d", dict(cells=(16, 20, None), dissipation=dict(mu=1e-3, kappa=1e-3)), 2),
    ("sod", dict(cells=(64, None, None), dissipation=dict(mu=2e-3, prandtl=0.7)), 2),
    # WENO5-JS
    ("sod", dict(cells=(80, None, None), stencil="WENO5-JS"), 2),
    ("tgv", dict(cells=(10, 10, 10), stencil="WENO5-JS", recon="PRIMITIVE", bc="PERIODIC"), 1),
    # HLLC signal speed estimates
    ("sod", dict(cells=(64, None, None), signal_speed="TORO"), 2),
    ("sod", dict(cells=(64, None, None), signal_speed="ARITHMETIC"), 2),
    ("riemann2d", dict(cells=(16, 16, None), signal_speed="DAVIS"), 1),
    ("riemann2d", dict(cells=(16, 16, None), signal_speed="RUSANOV"), 1),
    # HLL Riemann solver
    ("sod", dict(cells=(64, None, None), riemann="HLL"), 2),
    ("riemann2d", dict(cells=(16, 16, None), riemann="HLL", signal_speed="DAVIS"), 1),
    # positivity flux limiter (the shipped double-rarefaction example; a stronger one where it fires every step;
    # NASA / CELLSIZE variants)
    ("rarefaction", dict(cells=(100, None, None)), 12),
    ("rarefaction", dict(cells=(80, None, None), initial_condition={
        "u": "lambda x: -2.5*(x <= 0.5) + 2.5*(x > 0.5)", "p": 0.05}), 12),
    ("riemann2d", dict(cells=(16, 20, None), positivity={"flux_limiter": "NASA", "flux_partition": "CELLSIZE"}), 2),
    ("tgv", dict(cells=(8, 10, 12), positivity={"flux_limiter": "SIMPLE", "flux_partition": "CELLSIZE"}), 1),
    # the shipped lid-driven cavity / Rayleigh-Taylor / heat-equation examples (WALL, DIRICHLET, gravity, limiter,
    # WENO5-JS, viscous, heat flux only)
    ("cavity", dict(cells=(16, 14, None)), 2),
    ("rti", dict(cells=(12, 32, None)), 2),
    ("heat1d", dict(cells=(32, None, None)), 2),
    # RK2_LS4 and the generic reconstruction stencils
    ("sod", dict(cells=(64, None, None), integrator="RK2_LS4"), 2),
    ("tgv", dict(cells=(10, 8, 12), stencil="TENO5", integrator="RK2_LS4"), 1),
    ("tgv", dict(cells=(8, 8, 10), stencil="WENO6-CU", bc="PERIODIC"), 1),
    ("riemann2d", dict(cells=(16, 20, None), stencil="WENO3-JS", recon="PRIMITIVE"), 2),
    ("riemann2d", dict(cells=(16, 20, None), stencil="WENO3-Z"), 2),
    ("sod", dict(cells=(64, None, None), stencil="WENO1"), 2),
    ("sod", dict(cells=(64, None, None), stencil="KOREN"), 2),
    ("sod", dict(cells=(64, None, None), stencil="MC", recon="PRIMITIVE"), 2),
    ("riemann2d", dict(cells=(16, 16, None), stencil="MINMOD"), 1),
    ("riemann2d", dict(cells=(16, 16, None), stencil="SUPERBEE", recon="PRIMITIVE"), 1),
    ("sod", dict(cells=(64, None, None), stencil="VANALBADA"), 2),
    ("sod", dict(cells=(64, None, None), stencil="VANLEER", riemann="RUSANOV"), 2),
    ("sod", dict(cells=(64, None, None), stencil="WENO3-N"), 2),
    ("riemann2d", dict(cells=(16, 16, None), stencil="CENTRAL2", recon="PRIMITIVE"), 1),
    ("tgv", dict(cells=(8, 8, 10), stencil="TENO6"), 1),
    # convective_solver = FLUX-SPLITTING: the shipped Lax / Woodward-Colella examples and the other eigenvalue choices
    ("lax", dict(cells=(80, None, None)), 3),
    ("woodward", dict(cells=(100, None, None)), 3),
    ("sod", dict(cells=(64, None, None), flux_splitting="CLLF", stencil="WENO6-CU"), 2),
    ("riemann2d", dict(cells=(16, 20, None), flux_splitting="LLF"), 2),
    ("tgv", dict(cells=(10, 8, 12), flux_splitting="ROE"), 1),
    ("tgv", dict(cells=(8, 8, 10), flux_splitting="CLLF", bc="PERIODIC", stencil="TENO5"), 1),
    # reconstruction_variable CONSERVATIVE / CHAR-CONSERVATIVE, frozen_state ROE (godunov and flux_splitting blocks)
    ("sod", dict(cells=(64, None, None), recon="CONSERVATIVE"), 2),
    ("riemann2d", dict(cells=(16, 20, None), recon="CHAR-CONSERVATIVE", stencil="TENO5", riemann="HLL"), 2),
    ("tgv", dict(cells=(10, 8, 12), recon="CHAR-CONSERVATIVE"), 1),
    ("rarefaction", dict(cells=(80, None, None), recon="CHAR-CONSERVATIVE"), 6),
    ("sod", dict(cells=(64, None, None), frozen_state="ROE"), 2),
    ("tgv", dict(cells=(10, 8, 12), frozen_state="ROE", recon="CHAR-CONSERVATIVE"), 1),
    ("riemann2d", dict(cells=(16, 20, None), frozen_state="ROE", flux_splitting="CLLF"), 2),
    ("lax", dict(cells=(80, None, None), frozen_state="ROE"), 2),
    # the shipped 2-D heat equation example: DIRICHLET data given as a lambda of the transverse coordinate
    ("heat2d", dict(cells=(20, 16, None)), 3),
    # SIMPLE_INFLOW / SIMPLE_OUTFLOW / NEUMANN boundaries
    ("sod", dict(cells=(64, None, None), boundary_conditions={
        "west": {"type": "SIMPLE_INFLOW", "primitives_callable": {"rho": 1.0, "u": 0.3, "v": 0.0, "w": 0.0}},
        "east": {"type": "SIMPLE_OUTFLOW", "primitives_callable": {"p": 0.1}}}), 3),
    ("riemann2d", dict(cells=(16, 20, None), dissipation=dict(mu=1e-3, kappa=1e-3), boundary_conditions={
        "west": {"type": "SIMPLE_INFLOW", "primitives_callable": {"rho": "lambda y,t: 0.5 + 0.2 * y", "u": 1.2, "v": 0.0, "w": 0.0}},
        "east": {"type": "SIMPLE_OUTFLOW", "primitives_callable": {"p": "lambda y,t: 1.0 + 0.5 * y"}},
        "north": {"type": "NEUMANN", "primitives_callable": {"rho": 0.1, "u": "lambda x,t: 0.2 * x", "v": 0.0, "w": 0.0, "p": -0.3}},
        "south": {"type": "NEUMANN", "primitives_callable": {"rho": "lambda x,t: 0.3 * jnp.cos(5 * x)", "u": 0.0, "v": 0.1, "w": 0.0,
                                                             "p": 0.2}}}), 2),
    # the shipped double Mach reflection example: two boundary types on the south face
    ("dmr", dict(cells=(48, 32, None)), 3),
    ("sod", dict(cells=(64, None, None), stencil="TENO5-A"), 2),
    ("tgv", dict(cells=(8, 8, 10), stencil="TENO6-A"), 1),
    # WALL with a space-dependent wall velocity (a regularised lid on the shipped cavity)
    ("cavity", dict(cells=(16, 14, None), boundary_conditions={"north": {"type": "WALL", "wall_velocity_callable": {
        "u": "lambda x,t: 16.0 * x**2 * (1.0 - x)**2", "v": 0.0, "w": 0.0}}}), 2),
    # HLLC-LM and AUSM+
    ("sod", dict(cells=(64, None, None), riemann="HLLC-LM"), 3),
    ("tgv", dict(cells=(10, 8, 12), riemann="HLLC-LM"), 1),
    ("riemann2d", dict(cells=(16, 20, None), riemann="HLLC-LM", signal_speed="DAVIS", recon="PRIMITIVE"), 2),
    ("sod", dict(cells=(64, None, None), riemann="AUSMP"), 3),
    ("riemann2d", dict(cells=(16, 20, None), riemann="AUSMP", recon="PRIMITIVE"), 2),
    ("tgv", dict(cells=(8, 8, 10), riemann="AUSMP", bc="PERIODIC"), 1),
])
def test_port_is_bit_identical_to_reference(name, kw, nsteps):
    from oracle.refharness import pin_check
    kw = dict(kw)
    with np.errstate(all="ignore"):
        pin_check.check(name, nsteps=nsteps, **kw)
