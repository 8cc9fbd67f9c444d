"""Initial conditions for homogeneous isotropic turbulence (host NumPy, evaluated once at initialisation).

`initialize_hit` follows the reference's generator for `initial_condition/turbulent/case = "HIT"`
(turbulence/initialization/hit.py:20-240, :331-455, :471-533, :701-755; spectra :189-240; energy spectrum
turbulence/statistics/utilities/energy_spectrum.py:63-102; wavenumber grids math/fft/wavenumber.py:11-44, :126-130;
seeding turbulence/initialization/turb_init_manager.py:44-45) on ONE block holding the global grid: `np.random.seed`
+ `np.random.uniform` draws in the reference's order, `np.fft.rfftn / irfftn`, the same shell binning.  Variant IC1
(solenoidal velocity, uniform density and pressure) in both of the reference's forms -- physical-space random field
rescaled to the target spectrum and projected three times (default), or spectral-space construction
(`is_velocity_spectral`).  IC2-IC4 (Poisson-equation pressure / density) are not on this path.

`synthetic_solenoidal_ic` is the benchmark's own recipe for BASELINE config 5: random-phase Fourier
modes with E(k) ~ k^4 exp(-2 k^2 / k0^2), `np.random.default_rng(seed)`, projected to be divergence-free, rho = 1,
uniform p, rescaled to the target turbulent Mach number.  It is what `bench.py --workload hit` injects through
`InitializationManager.initialization(user_prime_init=...)`.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np

EPS = float(np.finfo(np.float64).eps)     # config/precision.py get_eps() in double precision


def real_wavenumber_grid(n: int) -> np.ndarray:
    """(3, N, N, N//2+1) integer wavenumber vectors, the last axis the real-FFT one (wavenumber.py:11-44)."""
    nf = n // 2 + 1
    k = np.fft.fftfreq(n, 1 / n).astype(int)
    k_real = np.arange(nf).astype(int)
    return np.array(np.meshgrid(k, k, k_real, indexing="ij"))


def factor_real(k_field: np.ndarray) -> np.ndarray:
    """Multiplicity of a real-FFT coefficient in the full spectrum (wavenumber.py:126-130)."""
    nyq = k_field.shape[1] // 2
    return 2 * (k_field[2] > 0) * (k_field[2] < nyq) + 1 * (k_field[2] == 0) + 1 * (k_field[2] == nyq)


def energy_spectrum_spectral(buffer_hat: np.ndarray, n: int, multiplicative_factor: float = 1.0) -> np.ndarray:
    """Shell-summed energy spectrum of a velocity field in spectral space (energy_spectrum.py:63-102)."""
    eps = 1e-10
    k_field = real_wavenumber_grid(n)
    k_mag_vec = np.arange(n)
    fact = factor_real(k_field)
    kmag = np.sqrt(np.sum(np.square(k_field), axis=0))
    shell = (kmag + 0.5).astype(int).flatten()
    buffer_hat = buffer_hat / n ** 3
    abs_energy = np.sum(np.real(buffer_hat * np.conj(buffer_hat)), axis=-4)
    abs_energy = abs_energy * (fact * multiplicative_factor)
    n_samples = np.zeros(n)
    np.add.at(n_samples, shell, fact.flatten())
    spec = np.zeros(n)
    np.add.at(spec, shell, abs_energy.flatten())
    return spec * (4 * np.pi * k_mag_vec * k_mag_vec / (n_samples + eps))


def get_target_spectrum(name: str) -> Callable:
    """hit.py:189-240."""
    name = name.upper()
    if name == "KOLMOGOROV":
        def ek(k, xi_0, xi_1, **kw):
            k = k + np.where(k == 0, EPS, 0)
            return (k >= xi_0) * (k < xi_1) * (0.5 * k ** (-5 / 3))
    elif name == "EXPONENTIAL":
        def ek(k, xi_0, u_rms, **kw):
            a = u_rms ** 2 * 16 * np.sqrt(2 / np.pi)
            return a * k ** 4 / xi_0 ** 5 * np.exp(-2 * k ** 2 / xi_0 ** 2)
    elif name == "BOX":
        def ek(k, xi_0, xi_1, **kw):
            return (k >= xi_0) * (k < xi_1) * 1.0
    else:
        raise NotImplementedError(f"energy_spectrum '{name}'")
    return ek


def rescale_field(velocity: np.ndarray, ek_target: np.ndarray) -> np.ndarray:
    """hit.py:701-755 (single block)."""
    n = velocity.shape[-1]
    k_field = real_wavenumber_grid(n)
    kmag2 = np.sum(np.square(k_field), axis=0)
    shell = (np.sqrt(kmag2 + EPS) + 0.5).astype(int)
    vhat = np.fft.rfftn(velocity, axes=(-3, -2, -1))
    ek_current = energy_spectrum_spectral(vhat, n, multiplicative_factor=0.5)
    scale = np.sqrt(ek_target / (ek_current + EPS))
    # the reference's `buffer_hat /= N**3` inside energy_spectrum_spectral acts on ITS argument in place under NumPy
    # semantics but not under JAX's: JAX arrays are immutable, so velocity_hat keeps its value there.  Restated as JAX.
    vhat = vhat * (scale[shell] * n ** 3)
    return np.fft.irfftn(vhat, axes=(-3, -2, -1))


def get_solenoidal_field(velocity: np.ndarray) -> np.ndarray:
    """Helmholtz projection in spectral space (hit.py:471-533, single block)."""
    n = velocity.shape[-1]
    k_field = real_wavenumber_grid(n)
    one_k2 = 1.0 / (np.sum(k_field * k_field, axis=0) + EPS)
    vhat = np.fft.rfftn(velocity, axes=(-3, -2, -1))
    div = np.sum(k_field * vhat, axis=0)
    return np.fft.irfftn(vhat - k_field * one_k2 * div, axes=(-3, -2, -1))


def _conjugate_symmetry_2d(a: np.ndarray) -> np.ndarray:
    """hit.py:303-327."""
    n1, n2 = a.shape[-2:]
    assert n1 == n2, "Only implemented for square matrices."
    nf = n1 // 2
    neg, pos = np.s_[-nf + 1:], np.s_[1:nf]
    a[..., neg, neg] = np.flip(np.conj(a[..., pos, pos]), axis=(-2, -1))
    a[..., pos, neg] = np.flip(np.conj(a[..., neg, pos]), axis=(-2, -1))
    a[..., 0, neg] = np.flip(np.conj(a[..., 0, pos]), axis=-1)
    a[..., neg, 0] = np.flip(np.conj(a[..., pos, 0]), axis=-1)
    return a


def solenoidal_velocity_spectral(n: int, gamma: float, R: float, T_ref: float, ek_fun: Callable, ma_target: float,
                                 xi_0: int, xi_1: int) -> np.ndarray:
    """Spectral-space construction after Johnsen et al. 2010 (hit.py:331-420, single block)."""
    c_ref = np.sqrt(gamma * R * T_ref)
    u_rms = ma_target / np.sqrt(3) * c_ref
    k_field = real_wavenumber_grid(n)
    k_mag = np.sqrt(np.sum(k_field * k_field, axis=0))
    k12 = np.sqrt(k_field[0] * k_field[0] + k_field[1] * k_field[1])
    k_mag[0, 0, 0] = EPS
    ek = ek_fun(k_mag, u_rms=u_rms, xi_0=xi_0, xi_1=xi_1)
    amplitude = np.sqrt(2 * ek / (4 * np.pi * k_mag * k_mag))
    phi = 2 * np.pi * np.random.uniform(size=k_field.shape)
    a = amplitude * np.exp(1j * phi[0]) * np.cos(phi[2])
    b = amplitude * np.exp(1j * phi[1]) * np.sin(phi[2])
    one_k = 1.0 / k_mag
    k1 = k_field[0] / (k12 + 1e-100)
    k2 = k_field[1] / (k12 + 1e-100)
    k2[0, 0, :] = 1.0
    vhat = np.array([k2 * a + k1 * k_field[2] * one_k * b,
                     k2 * k_field[2] * one_k * b - k1 * a,
                     -k12 * one_k * b], dtype=np.complex128)
    vhat[:, 0, 0, 0] = 0.0
    vhat[:, n // 2, :, :] = 0.0
    vhat[:, :, n // 2, :] = 0.0
    vhat[:, :, :, -1] = 0.0
    vhat[..., 0] = _conjugate_symmetry_2d(vhat[..., 0])
    vhat *= n ** 3
    return np.fft.irfftn(vhat, axes=(-3, -2, -1))


def initialize_hit(n: int, gamma: float, R: float, *, energy_spectrum: str, xi_0: int, ma_target: float, T_ref: float,
                   rho_ref: float, ic_type: str = "IC1", xi_1: int = 16, is_velocity_spectral: bool = False,
                   random_seed: Optional[int] = 0) -> np.ndarray:
    """(5, N, N, N) primitives of the reference's HIT initial condition on the global grid (hit.py:20-186).
    The caller seeds like the reference does (turb_init_manager.py:44-45) by passing `random_seed`."""
    if ic_type != "IC1":
        raise NotImplementedError(f"initial_condition/turbulent/parameters/ic_type '{ic_type}' is not implemented on "
                                  "the B200 path (implemented: IC1)")
    if random_seed is not None:
        np.random.seed(random_seed)
    p_ref = rho_ref * R * T_ref
    c_ref = np.sqrt(gamma * p_ref / rho_ref)
    ek_fun = get_target_spectrum(energy_spectrum)
    if is_velocity_spectral:
        assert energy_spectrum.upper() == "EXPONENTIAL", \
            "For velocity initialization in spectral space, choose exponential energy spectrum."
        velocity = solenoidal_velocity_spectral(n, gamma, R, T_ref, ek_fun, ma_target, xi_0, xi_1)
    else:
        ek_target = ek_fun(np.arange(n), xi_0=xi_0, xi_1=xi_1, u_rms=1.0)
        velocity = 2 * np.pi * np.random.uniform(size=(3, n, n, n))
        for _ in range(3):
            velocity = rescale_field(velocity, ek_target)
            velocity = get_solenoidal_field(velocity)
        q_rms = np.sqrt(np.mean(np.sum(velocity * velocity, axis=0)))
        velocity = velocity * (ma_target / (q_rms / c_ref))
    pressure = p_ref * np.ones_like(velocity[0])
    density = rho_ref * np.ones_like(velocity[0])
    return np.concatenate([density[None], velocity, pressure[None]], axis=0)


def synthetic_mode_table(k0: float = 4.0, seed: int = 0):
    """The Fourier modes of `synthetic_solenoidal_ic`: complex vector amplitudes c[q, kx, ky, kz] on the half space
    kz > 0 | (kz = 0, ky > 0) | (kz = ky = 0, kx > 0), |k| <= 3 k0 (beyond, E(k) < 1e-6 of its peak), each perpendicular
    to its wavevector, |c_k|^2 ~ E(|k|) / (4 pi k^2) with E(k) = k^4 exp(-2 k^2 / k0^2), random phase and direction from
    `np.random.default_rng(seed)`.  Returns (c (3, 2K+1, 2K+1, K+1) complex, K, mean of u.u of the field it sums to)."""
    K = int(np.ceil(3 * k0))
    rng = np.random.default_rng(seed)
    kx, ky, kz = np.meshgrid(np.arange(-K, K + 1), np.arange(-K, K + 1), np.arange(0, K + 1), indexing="ij")
    k2 = (kx * kx + ky * ky + kz * kz).astype(float)
    half = (kz > 0) | ((kz == 0) & (ky > 0)) | ((kz == 0) & (ky == 0) & (kx > 0))
    on = half & (k2 <= K * K)
    kmag = np.sqrt(np.where(on, k2, 1.0))
    amp = np.where(on, np.sqrt(kmag ** 4 * np.exp(-2 * kmag ** 2 / k0 ** 2) / (4 * np.pi * kmag ** 2)), 0.0)
    phase = rng.uniform(0, 2 * np.pi, size=on.shape)
    r = rng.normal(size=(3,) + on.shape)
    kv = np.stack([kx, ky, kz]).astype(float)
    perp = np.cross(r, kv, axis=0)
    norm = np.sqrt((perp ** 2).sum(axis=0))
    perp = perp / np.where(norm > 0, norm, 1.0)
    c = (amp * np.exp(1j * phase))[None] * perp
    mean_uu = 0.5 * float(np.sum(np.abs(c) ** 2))        # <Re(c e^{ik.x}) . Re(c e^{ik.x})> = |c|^2 / 2 per mode
    return c, K, mean_uu


def synthetic_solenoidal_ic(n: int, *, gamma: float = 1.4, k0: float = 4.0, ma_t: float = 0.4, seed: int = 0,
                            block: Optional[Tuple[slice, slice, slice]] = None, device=None, out=None):
    """BASELINE config 5 / SURVEY 8(d): periodic [0, 2 pi]^3, rho = 1, uniform p = 1 / gamma (so c = 1), and a
    divergence-free velocity field of random-phase Fourier modes with E(k) ~ k^4 exp(-2 k^2 / k0^2) scaled to the
    turbulent Mach number q_rms / c = ma_t:

        u(x) = Re sum_k c_k exp(i k.x),   c_k . k = 0,   k from `synthetic_mode_table` (a few thousand modes).

    The field is a smooth band-limited FUNCTION, evaluated at the cell centres of any `block` (slices of the global
    cell-index ranges) of any grid with three separable contractions (kz, then ky, then kx -- small GEMMs), so every
    rank builds its own block without a global FFT or a global array: 1024^3 needs no 8.6 GB spectral buffers.
    `device`: a torch device to run the contractions on (the GPU at bench sizes); default NumPy.  `out`: optional
    (5, bx, by, bz) array / tensor view to fill (e.g. the interior of a halo'd buffer).  Returns (5, bx, by, bz)."""
    c, K, mean_uu = synthetic_mode_table(k0, seed)
    scale = ma_t / np.sqrt(mean_uu)
    dx = 2 * np.pi / n
    xs = (np.arange(n) + 0.5) * dx
    sl = block if block is not None else (slice(None),) * 3
    x, y, z = xs[sl[0]], xs[sl[1]], xs[sl[2]]
    kk = np.arange(-K, K + 1)
    ex = np.exp(1j * np.outer(x, kk))                      # (bx, 2K+1)
    ey = np.exp(1j * np.outer(kk, y))                      # (2K+1, by)
    ez = np.exp(1j * np.outer(np.arange(0, K + 1), z))     # (K+1, bz)
    shape = (len(x), len(y), len(z))
    if device is None:
        res = np.empty((5,) + shape) if out is None else out
        for q in range(3):
            f = c[q] @ ez                                  # (2K+1, 2K+1, bz): sum over kz
            g = np.einsum("aby,abz->ayz", ey[None].repeat(2 * K + 1, 0), f)     # sum over ky -> (2K+1, by, bz)
            g = g.reshape(2 * K + 1, -1)
            u = ex.real @ g.real - ex.imag @ g.imag        # sum over kx, real part -> (bx, by * bz)
            res[1 + q] = (scale * u).reshape(shape)
        res[0] = 1.0
        res[4] = 1.0 / gamma
        return res
    import torch
    dev = torch.device(device)
    res = torch.empty((5,) + shape, dtype=torch.float64, device=dev) if out is None else out
    t = lambda a: torch.as_tensor(a, device=dev)
    ex_t, ey_t, ez_t = t(ex), t(ey), t(ez)
    for q in range(3):
        f = torch.matmul(t(c[q]), ez_t)                                    # (2K+1, 2K+1, bz)
        g = torch.einsum("by,abz->ayz", ey_t, f).reshape(2 * K + 1, -1)    # (2K+1, by * bz)
        # x in slabs: bounds the temporaries to a few hundred MB at 1024^3
        gr, gi = g.real.contiguous(), g.imag.contiguous()
        step = max(1, (1 << 27) // max(1, shape[1] * shape[2]))
        for i0 in range(0, shape[0], step):
            i1 = min(shape[0], i0 + step)
            u = ex_t.real[i0:i1] @ gr - ex_t.imag[i0:i1] @ gi
            res[1 + q, i0:i1] = (scale * u).reshape(i1 - i0, shape[1], shape[2])
    res[0] = 1.0
    res[4] = 1.0 / gamma
    return res
