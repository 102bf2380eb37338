"""Synthetic atmospheres for the BASELINE.json configurations, numpy only.

All arrays use the layouts that cross the reference's C ABI (cpp/include/c_api/atmosphere.h:86-91):
``ssa``/``total_extinction`` are Fortran-ordered ``[nloc, nwavel]``, ``leg_coeff`` is Fortran-ordered
``[nleg, nloc, nwavel]``.  Seeds are fixed, so every rank / test / bench run sees the same inputs.

The Rayleigh-like extinction profile is analytic (exp scale height), standing in for the 101-level table of
``src/sasktran2/test_util/scenarios.py:17-119`` — same magnitude (7e-5 m^-1 at the ground, ~0.5 vertical OD).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class Scenario:
    name: str
    nstr: int
    altitudes: np.ndarray            # [nloc] metres, ascending
    interp: int                      # 0 shell, 1 linear, 2 lower
    geotype: int                     # 0 plane-parallel, 1 pseudo-spherical
    cos_sza: float
    los_cos_vza: np.ndarray          # [nlos]
    los_rel_az: np.ndarray           # [nlos]
    observer_altitude: float
    ssa: np.ndarray                  # [nloc, nw] F
    total_extinction: np.ndarray     # [nloc, nw] F
    leg_coeff: np.ndarray            # [nleg, nloc, nw] F
    albedo: np.ndarray               # [nw]
    solar_irradiance: np.ndarray     # [nw]
    earth_radius: float = 6372000.0
    # weighting-function mappings: name -> dict(d_extinction, d_ssa [nloc,nw] F, optional d_legendre
    # [nleg,nloc,nw] F + scat_factor [nloc,nw] F)
    mappings: dict = field(default_factory=dict)
    # spherical (limb) scenarios: rays in the oracle's notation, ("tangent", tangent_altitude, rel_az, observer_altitude,
    # cos_sza) / ("ground", cos_sza, rel_az, cos_vza, observer_altitude); SZAs of the DO source table; solar azimuth
    rays: list = field(default_factory=list)
    num_sza: int = 1
    saa: float = 0.0

    @property
    def nwavel(self) -> int:
        return self.ssa.shape[1]

    @property
    def nloc(self) -> int:
        return self.altitudes.size

    @property
    def nlos(self) -> int:
        return len(self.rays) if self.rays else self.los_cos_vza.size


# Grey continuum absorption as a fraction of the Rayleigh extinction.  It keeps every layer's single-scatter
# albedo at least ~1e-3 away from 1: in (nearly) conservative layers the DO eigen-solution loses digits by
# cancellation in S+ X (noise ~ eps / (1 - omega), ~1e-7 at the reference's 1e-9 dither, for the reference
# and for this implementation alike — DESIGN.md "Conditioning"), which would turn a 1e-9 parity check into a
# comparison of rounding noise.
CONTINUUM_ABSORPTION = 1.0e-3


def rayleigh_extinction(z_m: np.ndarray) -> np.ndarray:
    return 7.0e-5 * np.exp(-z_m / 7400.0)


def _mix(components, nleg):
    """components: list of (k [nloc,nw], omega [nloc,nw] or scalar, beta [nleg] or [nleg,nloc,nw])."""
    k_tot = sum(c[0] for c in components)
    ks_tot = sum(c[0] * c[1] for c in components)
    leg = np.zeros((nleg,) + k_tot.shape)
    for k, w, b in components:
        b = np.asarray(b, dtype=float)
        if b.ndim == 1:
            b = b[:, None, None]
        leg += (k * w)[None] * b
    with np.errstate(invalid="ignore", divide="ignore"):
        leg = np.where(ks_tot[None] > 0, leg / ks_tot[None], 0.0)
        ssa = np.where(k_tot > 0, ks_tot / k_tot, 0.0)
    leg[0] = np.where(ks_tot > 0, 1.0, leg[0])
    return k_tot, ssa, leg, ks_tot


def hg_moments(g: float, nleg: int) -> np.ndarray:
    l = np.arange(nleg)
    return (2 * l + 1) * g**l


def rayleigh_moments(nleg: int) -> np.ndarray:
    b = np.zeros(nleg)
    b[0] = 1.0
    if nleg > 2:
        b[2] = 0.5
    return b


def config1(nwavel: int = 1000, nlayers: int = 50) -> Scenario:
    """C1: plane-parallel DO, 4 streams, 50 layers, Rayleigh + O3, 1 nadir LOS."""
    nleg = 4
    z = np.linspace(0.0, 100e3, nlayers + 1)
    s = np.logspace(-1, 1, nwavel)
    k_ray = rayleigh_extinction(z)[:, None] * s[None, :]
    g_lam = 0.5 * (1 - np.cos(2 * np.pi * np.arange(nwavel) / max(nwavel - 1, 1) * 3.0))
    # keep a small absorption floor so no layer is exactly conservative
    k_o3 = 3e-5 * np.exp(-(((z - 25e3) / 8e3) ** 2))[:, None] * (0.02 + g_lam)[None, :]
    k, ssa, leg, _ = _mix([(k_ray, 1.0, rayleigh_moments(nleg)), (k_o3, 0.0, np.zeros(nleg)),
                           (CONTINUUM_ABSORPTION * k_ray, 0.0, np.zeros(nleg))], nleg)
    return Scenario("C1", 4, z, 1, 0, 0.6, np.array([1.0]), np.array([0.0]), 200e3,
                    np.asfortranarray(ssa), np.asfortranarray(k), np.asfortranarray(leg),
                    np.full(nwavel, 0.3), np.ones(nwavel))


def config2(nwavel: int = 100000, nlayers: int = 100, nstr: int = 16, nlos: int = 10, with_wf: bool = False,
            seed: int = 0, block: tuple[int, int] | None = None, nleg: int | None = None) -> Scenario:
    """C2 (and C5 when with_wf): pseudo-spherical DO, 16 streams, 100 layers, Rayleigh + aerosol (+ O3/NO2
    absorbers so the atmosphere is not conservative), 10 ground-viewing LOS.

    `block = (start, count)` builds only wavelengths [start, start + count) of the `nwavel`-point spectrum (what one
    rank of a wavelength-sharded run owns); the arrays equal the corresponding slices of the full scenario.
    `nleg` > nstr stores more phase moments than streams (what delta-M scaling needs)."""
    nleg = nstr if nleg is None else nleg
    z = np.linspace(0.0, 100e3, nlayers + 1)
    widx = np.arange(nwavel) if block is None else np.arange(block[0], block[0] + block[1])
    lam_frac = widx / max(nwavel - 1, 1)
    s = 10.0 ** (-1.0 + 2.0 * lam_frac) if nwavel > 1 else np.array([0.1])
    nwavel = widx.size
    k_ray = rayleigh_extinction(z)[:, None] * s[None, :]
    k_aer = (1e-5 * np.exp(-z / 3e3))[:, None] * np.ones(nwavel)[None, :]
    n_air = np.exp(-z / 7400.0)
    o3_shape = np.exp(-(((z - 25e3) / 8e3) ** 2))
    sig_o3 = 3e-6 * (0.05 + 0.5 * (1 - np.cos(2 * np.pi * 5.0 * lam_frac)))
    sig_no2 = 4e-7 * (0.05 + 0.5 * (1 + np.sin(2 * np.pi * 11.0 * lam_frac)))
    vmr_o3 = o3_shape / n_air.clip(1e-6)  # so that vmr * n_air = o3_shape
    vmr_o3 = np.minimum(vmr_o3, 50.0)
    vmr_no2 = np.exp(-(((z - 30e3) / 10e3) ** 2))
    k_o3 = (vmr_o3 * n_air)[:, None] * sig_o3[None, :]
    k_no2 = (vmr_no2 * n_air)[:, None] * sig_no2[None, :]
    w_aer = 0.95
    b_aer = hg_moments(0.7, nleg)
    comps = [(k_ray, 1.0, rayleigh_moments(nleg)), (k_aer, w_aer, b_aer), (k_o3, 0.0, np.zeros(nleg)),
             (k_no2, 0.0, np.zeros(nleg)), (CONTINUUM_ABSORPTION * k_ray, 0.0, np.zeros(nleg))]
    k, ssa, leg, ks = _mix(comps, nleg)
    sc = Scenario("C5" if with_wf else "C2", nstr, z, 1, 1, 0.6, np.linspace(1.0, 0.55, nlos),
                  np.linspace(0.0, np.pi, nlos), 200e3, np.asfortranarray(ssa), np.asfortranarray(k),
                  np.asfortranarray(leg), np.full(nwavel, 0.3), np.ones(nwavel))
    if with_wf:
        # Absorber VMR mappings (rust/sasktran2-rs/src/constituent/types/vmr_alt_absorber.rs:380-398):
        #   d_extinction = sigma * n_air, d_ssa = -omega * d_extinction / k
        for name, sig in (("wf_o3_vmr", sig_o3), ("wf_no2_vmr", sig_no2)):
            d_ext = n_air[:, None] * sig[None, :]
            d_ssa = -ssa * d_ext / k
            sc.mappings[name] = dict(d_extinction=np.asfortranarray(d_ext), d_ssa=np.asfortranarray(d_ssa))
        # Scatterer extinction mapping (per unit aerosol extinction at each grid point):
        #   d_extinction = 1, d_ssa = (w_aer - omega)/k, d_legendre = beta_aer - beta_mix, scat_factor = w_aer/(omega k)
        d_ext = np.ones_like(k)
        d_ssa = (w_aer - ssa) / k
        d_leg = b_aer[:, None, None] - leg
        scat_factor = w_aer / ks
        sc.mappings["wf_aerosol_extinction"] = dict(d_extinction=np.asfortranarray(d_ext), d_ssa=np.asfortranarray(d_ssa),
                                                    d_legendre=np.asfortranarray(d_leg),
                                                    scat_factor=np.asfortranarray(scat_factor))
    return sc


def config4(nwavel: int = 10000, nlayers: int = 100, nstr: int = 16, nrays: int = 100, num_sza: int = 2,
            block: tuple[int, int] | None = None) -> Scenario:
    """C4: OSIRIS-style limb scan in spherical geometry: exact single scatter + discrete-ordinates multiple-scatter
    source, `nrays` TangentAltitudeSolar lines of sight (tangent altitudes linspace(10, 60 km), relative azimuth 0.3,
    observer at 200 km, cos_sza 0.6), `num_sza` SZAs for the DO source table, the C2 atmosphere."""
    sc = config2(nwavel=nwavel, nlayers=nlayers, nstr=nstr, nlos=1, with_wf=False, block=block)
    sc.name = "C4"
    sc.geotype = 2
    sc.los_cos_vza = np.zeros(0)
    sc.los_rel_az = np.zeros(0)
    sc.rays = [("tangent", float(h), 0.3, 200e3, sc.cos_sza) for h in np.linspace(10e3, 60e3, nrays)]
    sc.num_sza = num_sza
    return sc


def config3(nwavel: int = 1000000, nlayers: int = 60, nlos: int = 2, block: tuple[int, int] | None = None,
            seed: int = 0) -> Scenario:
    """C3: two-stream source, 60 layers, line-by-line O2-A-band-like spectrum: k = k_ray + k_line with
    k_line = 10^U(-9, -3) exp(-z / 8 km) per wavelength (a counter-based hash of the wavelength index, so that a
    block of the spectrum equals the slice of the full one), beta = [1, 0.1, 0.5], plane parallel, ground-viewing LOS
    (SURVEY.md section 8d).  `block = (start, count)` builds only that part of the spectrum."""
    nleg = 3
    z = np.linspace(0.0, 60e3, nlayers + 1)
    widx = np.arange(nwavel, dtype=np.uint64) if block is None else np.arange(block[0], block[0] + block[1], dtype=np.uint64)
    # splitmix64 of (index, seed) -> uniform [0, 1)
    x = widx + np.uint64((0x9E3779B97F4A7C15 * (seed + 1)) % (1 << 64))   # wraps like the 64-bit multiply
    with np.errstate(over="ignore"):
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    u = (x >> np.uint64(11)).astype(np.float64) / float(1 << 53)
    k_line = (10.0 ** (-9.0 + 6.0 * u))[None, :] * np.exp(-z / 8e3)[:, None]
    k_ray = (0.02 * rayleigh_extinction(z))[:, None] * np.ones(widx.size)[None, :]
    k = k_ray + k_line
    ssa = k_ray / k
    leg = np.zeros((nleg, z.size, widx.size))
    leg[0], leg[1], leg[2] = 1.0, 0.1, 0.5
    return Scenario("C3", 2, z, 1, 0, 0.6, np.linspace(1.0, 0.8, nlos), np.linspace(0.0, 1.0, nlos), 200e3,
                    np.asfortranarray(ssa), np.asfortranarray(k), np.asfortranarray(leg), np.full(widx.size, 0.3),
                    np.ones(widx.size))


def small_wf_case(nstr: int = 8, nlayers: int = 12, nwavel: int = 3, nlos: int = 3, interp: int = 1, geotype: int = 1,
                  seed: int = 1, nleg: int | None = None) -> Scenario:
    """Small pseudo-spherical Rayleigh + aerosol + absorber case with all three mapping kinds; used by the
    weighting-function tests (finite differences in the spirit of src/sasktran2/test_util/wf.py:9-80)."""
    sc = config2(nwavel=nwavel, nlayers=nlayers, nstr=nstr, nlos=nlos, with_wf=True, seed=seed, nleg=nleg)
    sc.name = "small_wf"
    sc.interp = interp
    sc.geotype = geotype
    if nlos > 1:
        sc.los_cos_vza = np.linspace(0.95, 0.45, nlos)
    return sc
