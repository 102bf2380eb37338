"""Pins the pseudo-spherical chapman factors of the oracle (the headline configs are pseudo-spherical, the reference's
own golden tables are all plane-parallel): the restatement of the reference's ray-traced factors
(GeometryLayerArray::calculate_chapman_factors_raytracer, sktran_do_geometrylayerarray.cpp:122-186, over
SphericalShellRayTracer, cpp/lib/raytracing/spherical_shell.cpp) against (1) the closed straight-line formula of
calculate_chapman_factors (:69-119) that the CUDA host plan uses and (2) an independent computation: the path length of
the sun ray inside every spherical shell from the roots of |p + s e|^2 = r^2 in 80-digit arithmetic."""
import numpy as np
import pytest


def _independent_chapman(alt, cos_sza, earth_radius):
    import mpmath as mp

    mp.mp.dps = 80
    nl = len(alt) - 1
    out = np.zeros((nl, nl))
    csz = mp.mpf(cos_sza)
    e = (mp.sqrt(1 - csz * csz), mp.mpf(0), csz)
    for p in range(nl):                       # layer p: floor alt[nl - 1 - p]
        r0 = mp.mpf(earth_radius) + mp.mpf(float(alt[nl - 1 - p]))
        # |(0,0,r0) + s e|^2 = r^2  ->  s = -r0 csz + sqrt(r^2 - r0^2 (1 - csz^2))   (outward root)
        dist = lambda r: -r0 * csz + mp.sqrt(r * r - r0 * r0 * (1 - csz * csz))
        for q in range(p + 1):
            lo = mp.mpf(earth_radius) + mp.mpf(float(alt[nl - 1 - q]))
            hi = mp.mpf(earth_radius) + mp.mpf(float(alt[nl - q]))
            out[p, q] = float((dist(hi) - dist(lo)) / (hi - lo))
    return out


@pytest.mark.parametrize("cos_sza", [0.6, 0.25, 0.05, 0.999])
@pytest.mark.parametrize("grid", ["uniform-1km", "irregular"])
def test_ray_traced_chapman_factors(oracle_mod, cos_sza, grid):
    if grid == "uniform-1km":
        alt = np.linspace(0.0, 100e3, 101)
    else:
        alt = np.sort(np.concatenate([[0.0], np.random.default_rng(5).uniform(10.0, 80e3, 30)]))
    kw = dict(nstr=4, alt=alt, interp=1, geotype=1, cos_sza=cos_sza, earth_radius=6372000.0, los_cos_vza=[1.0],
              los_rel_az=[0.0])
    traced = oracle_mod.plan(**kw)["chapman"]
    formula = oracle_mod.plan(**kw, chapman_straight_line=True)["chapman"]
    exact = _independent_chapman(alt, cos_sza, 6372000.0)
    nz = exact != 0
    assert np.array_equal(traced != 0, nz) and np.array_equal(formula != 0, nz)   # lower triangle, nothing else
    # rounding of the shell intersections relative to the shell thickness: <= 1e-9 for shells >= 10 m
    thick = np.diff(alt)[::-1]
    tol = 1e-9 * np.maximum(1.0, 1e3 / thick)[None, :]
    assert np.all(np.abs(traced / np.where(nz, exact, 1) - 1)[nz] <= np.broadcast_to(tol, exact.shape)[nz])
    assert np.all(np.abs(formula / np.where(nz, exact, 1) - 1)[nz] <= np.broadcast_to(tol, exact.shape)[nz])
    if grid == "uniform-1km":
        np.testing.assert_allclose(traced, formula, rtol=1e-12, atol=0)


def test_pseudo_spherical_radiance_insensitive_to_chapman_evaluation(oracle_mod):
    """Radiances with the ray-traced and the closed-formula chapman factors agree far below the 1e-9 parity tolerance
    at the benchmark shape (the CUDA plan uses the closed formula, the oracle default is the ray tracer)."""
    from sasktran2_b200 import scenarios

    sc = scenarios.config2(nwavel=4, nlayers=100, nstr=8, nlos=3)
    kw = dict(nstr=sc.nstr, alt=sc.altitudes, interp=sc.interp, geotype=sc.geotype, cos_sza=sc.cos_sza,
              earth_radius=sc.earth_radius, los_cos_vza=sc.los_cos_vza, los_rel_az=sc.los_rel_az, ssa=sc.ssa,
              ext=sc.total_extinction, leg=sc.leg_coeff, albedo=sc.albedo)
    a = oracle_mod.do_radiance(**kw)["radiance"]
    oracle_mod.lib().oracle_set_chapman_straight_line(1)
    try:
        b = oracle_mod.do_radiance(**kw)["radiance"]
    finally:
        oracle_mod.lib().oracle_set_chapman_straight_line(0)
    np.testing.assert_allclose(a, b, rtol=1e-12)
