"""Spherical ("limb") line-of-sight path of the product (BASELINE config 4) against the limb
oracle (oracle/limb_oracle.hpp, itself pinned to the reference's golden numbers by tests/test_oracle_limb.py).

CPU part: the host-side geometry plan (ray tracing, optical-depth stencils, solar rays) through the C ABI.
GPU part: radiances and line-of-sight optical depths through Engine.calculate_radiance."""
import ctypes as C

import numpy as np
import pytest

import sasktran2_b200 as sk
from oracle import oracle
from sasktran2_b200 import _lib
from tests.test_oracle_limb import GOLDEN_OPTICAL_DEPTH, GOLDEN_RADIANCE, reference_case


def _geometry_and_view(c, interp=sk.InterpolationMethod.LinearInterpolation):
    geo = sk.Geometry1D(c["cos_sza"], c["saa"], c["earth_radius"], c["alt"], interp, sk.GeometryType.Spherical)
    view = sk.ViewingGeometry()
    for r in c["rays"]:
        if r[0] == "ground":
            view.add_ray(sk.GroundViewingSolar(r[1], r[2], r[3], r[4]))
        else:
            view.add_ray(sk.TangentAltitudeSolar(r[1], r[2], r[3], r[4]))
    return geo, view


@pytest.mark.parametrize("interp", [0, 1, 2])
def test_limb_plan_matches_oracle_geometry(interp):
    c = reference_case()
    c["interp"] = interp
    geo, view = _geometry_and_view(c, sk.InterpolationMethod(interp))
    assert view.num_rays == 4
    ext = np.ascontiguousarray(c["ext"][:, 1])
    out = np.zeros((4, 6))
    npts = C.c_int(0)
    sza = np.zeros(2)
    _lib.check(_lib.lib().sk_b200_limb_plan_check(geo._geometry, view._viewing_geometry, 8, 2, _lib.dptr(ext), _lib.dptr(out),
                                                  C.byref(npts), _lib.dptr(sza)), "limb_plan_check")
    ora = oracle.limb_radiance(**c, ms_do=False, exact_tangent=True)
    g = oracle.limb_geometry(alt=c["alt"], interp=interp, cos_sza=c["cos_sza"], saa=c["saa"], rays=c["rays"])
    np.testing.assert_allclose(out[:, 0], ora["los_optical_depth"][1], rtol=1e-9)
    # against the reference's tangent-layer arithmetic: its own cross-platform tolerance (see limb_oracle.hpp, exact_tangent_ref)
    np.testing.assert_allclose(out[:, 0], oracle.limb_radiance(**c, ms_do=False)["los_optical_depth"][1], rtol=5e-7)
    np.testing.assert_array_equal(out[:, 1].astype(int), g["nlayers"])
    np.testing.assert_allclose(out[:, 2], g["nlayers"], rtol=1e-12)      # interpolation weights of every segment sum to one
    np.testing.assert_allclose(out[:, 3], g["cos_scatter"], rtol=1e-12)
    assert npts.value > 0 and sza[0] < sza[1]
    # the sun is above the horizon of every point of these rays: no solar ray is blocked, the far end of a limb ray sits
    # at the top of the atmosphere (no optical depth towards the sun)
    assert np.all(out[:, 4:] >= 0.0)
    assert np.all(out[2:, 4] < 1e-12) and np.all(out[:, 5] < 1e-12)


def test_tangent_ray_refused_outside_spherical_geometry():
    c = reference_case()
    geo = sk.Geometry1D(c["cos_sza"], 0.0, c["earth_radius"], c["alt"], sk.InterpolationMethod.LinearInterpolation,
                        sk.GeometryType.PseudoSpherical)
    view = sk.ViewingGeometry()
    view.add_ray(sk.TangentAltitudeSolar(12_345.0, 0.0, 200_000.0, c["cos_sza"]))
    cfg = sk.Config()
    cfg.multiple_scatter_source = sk.MultipleScatterSource.DiscreteOrdinates
    cfg.single_scatter_source = sk.SingleScatterSource.DiscreteOrdinates
    with pytest.raises(_lib.SasktranError):
        sk.Engine(cfg, geo, view)


def _run(c, ms, ss, num_sza=2, interp=1, nstr=None, moments=16):
    cfg = sk.Config()
    cfg.num_streams = nstr or c["nstr"]
    cfg.num_stokes = 1
    cfg.num_sza = num_sza
    cfg.num_singlescatter_moments = moments
    cfg.multiple_scatter_source = ms
    cfg.single_scatter_source = ss
    cfg.output_los_optical_depth = True
    geo, view = _geometry_and_view(c, sk.InterpolationMethod(interp))
    eng = sk.Engine(cfg, geo, view)
    nw = c["ssa"].shape[1]
    atm = sk.Atmosphere(geo, cfg, numwavel=nw, calculate_derivatives=False, num_legendre=c["leg"].shape[0])
    atm.storage.ssa[:] = c["ssa"]
    atm.storage.total_extinction[:] = c["ext"]
    atm.storage.leg_coeff[:] = c["leg"]
    atm.surface.albedo[:] = c["albedo"]
    return eng.calculate_radiance(atm)


@pytest.mark.gpu
def test_cuda_limb_reference_regression_case():
    """The reference's own spherical golden case (tests/engine/test_1d_solver_regression.py:112-239): 5e-7 against its
    stored numbers, 1e-9 against the oracle."""
    c = reference_case()
    res = _run(c, sk.MultipleScatterSource.DiscreteOrdinates, sk.SingleScatterSource.DiscreteOrdinates)
    rad = res["radiance"][:, :, 0]
    np.testing.assert_allclose(rad, GOLDEN_RADIANCE, rtol=5e-7, atol=2e-13)
    np.testing.assert_allclose(res["los_optical_depth"], GOLDEN_OPTICAL_DEPTH, rtol=5e-7, atol=1e-13)
    ora = oracle.limb_radiance(**c, ms_do=True, ss_exact=False, exact_tangent=True)
    np.testing.assert_allclose(rad, ora["radiance"], rtol=1e-9)
    np.testing.assert_allclose(res["los_optical_depth"], ora["los_optical_depth"], rtol=1e-9)
    ora = oracle.limb_radiance(**c, ms_do=True, ss_exact=False)   # the reference's tangent-layer arithmetic
    np.testing.assert_allclose(rad, ora["radiance"], rtol=5e-7)


def limb_case(nstr=16, nlayers=30, nwavel=8, nrays=6, seed=0):
    """Small config-4 shaped case: Rayleigh + Henyey-Greenstein aerosol, tangent-altitude rays + one ground-viewing ray."""
    rng = np.random.default_rng(seed)
    alt = np.linspace(0.0, 60_000.0, nlayers + 1)
    s = np.logspace(-0.5, 0.5, nwavel)[None, :]
    k_ray = 1.2e-5 * np.exp(-alt / 8_000.0)[:, None] * s
    k_aer = 4e-6 * np.exp(-alt / 3_000.0)[:, None] * np.ones_like(s)
    k_abs = 2e-6 * np.exp(-((alt - 25_000.0) / 8_000.0) ** 2)[:, None] * rng.uniform(0.2, 1.0, (1, nwavel))
    ext = k_ray + k_aer + k_abs
    ssa = (k_ray + 0.95 * k_aer) / ext
    nleg = 16
    leg = np.zeros((nleg, alt.size, nwavel))
    b_ray = np.zeros(nleg)
    b_ray[0], b_ray[2] = 1.0, 0.5
    b_aer = (2 * np.arange(nleg) + 1) * 0.7 ** np.arange(nleg)
    f = (0.95 * k_aer) / (k_ray + 0.95 * k_aer)
    leg[:] = b_ray[:, None, None] * (1 - f)[None] + b_aer[:, None, None] * f[None]
    cos_sza = 0.6
    rays = [("tangent", float(h), 0.3, 200_000.0, cos_sza) for h in np.linspace(10_000.0, 50_000.0, nrays - 1)]
    rays.append(("ground", cos_sza, 0.8, 0.7, 200_000.0))
    return dict(nstr=nstr, alt=alt, interp=1, cos_sza=cos_sza, saa=0.0, earth_radius=6_372_000.0, rays=rays, num_sza=2,
                ssa=ssa, ext=ext, leg=leg, albedo=np.linspace(0.1, 0.5, nwavel))


@pytest.mark.gpu
@pytest.mark.parametrize("nstr,num_sza", [(16, 2), (8, 1), (4, 3), (2, 2)])
def test_cuda_limb_config4_shape_vs_oracle(nstr, num_sza):
    """Exact single scatter + DO multiple scatter (BASELINE config 4 shape): 1e-9 on every radiance."""
    c = limb_case(nstr=nstr)
    c["num_sza"] = num_sza
    res = _run(c, sk.MultipleScatterSource.DiscreteOrdinates, sk.SingleScatterSource.Exact, num_sza=num_sza)
    ora = oracle.limb_radiance(**c, ms_do=True, ss_exact=True, exact_tangent=True)
    err = np.max(np.abs(res["radiance"][:, :, 0] / ora["radiance"] - 1))
    ref = oracle.limb_radiance(**c, ms_do=True, ss_exact=True)   # the reference's tangent-layer arithmetic
    err_ref = np.max(np.abs(res["radiance"][:, :, 0] / ref["radiance"] - 1))
    print(f"limb nstr={nstr} num_sza={num_sza}: max rel diff vs oracle {err:.2e} (reference tangent arithmetic: {err_ref:.2e})")
    assert err < 1e-9
    assert err_ref < 5e-7   # the reference's own cross-platform tolerance for traced paths (test_1d_solver_regression.py:190-199)
    np.testing.assert_allclose(res["los_optical_depth"], ora["los_optical_depth"], rtol=1e-10)


@pytest.mark.gpu
@pytest.mark.parametrize("interp", [0, 2])
def test_cuda_limb_single_scatter_only_interpolation_modes(interp):
    c = limb_case(nstr=4, nwavel=4)
    c["interp"] = interp
    res = _run(c, sk.MultipleScatterSource.NoSource, sk.SingleScatterSource.Exact, interp=interp)
    ora = oracle.limb_radiance(**c, ms_do=False, ss_exact=True, exact_tangent=True)
    np.testing.assert_allclose(res["radiance"][:, :, 0], ora["radiance"], rtol=1e-10)


@pytest.mark.gpu
def test_cuda_limb_chunked_equals_unchunked():
    """Property at a larger spectrum: wavelength chunks of the limb path are independent."""
    c = limb_case(nstr=8, nwavel=96)
    a = _run(c, sk.MultipleScatterSource.DiscreteOrdinates, sk.SingleScatterSource.Exact)["radiance"].copy()
    import os
    os.environ["SK_B200_WORKSPACE_GB"] = "0.02"
    try:
        b = _run(c, sk.MultipleScatterSource.DiscreteOrdinates, sk.SingleScatterSource.Exact)["radiance"].copy()
    finally:
        del os.environ["SK_B200_WORKSPACE_GB"]
    np.testing.assert_array_equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("observer_altitude", [200_000.0, 700_000.0, 35_786_000.0])
def test_cuda_limb_vertical_rays_and_distant_observers(observer_altitude):
    """An exactly vertical ground-viewing ray (every point at one SZA: the SZA grid of the source table degenerates to
    LinSpaced(n, a, a), on which the reference divides by a zero spacing), a nearly vertical one, backward azimuth, a
    grazing tangent ray; observers from low orbit to geostationary."""
    c = limb_case(nstr=8, nlayers=20, nwavel=3, nrays=2, seed=3)
    o = observer_altitude
    c["rays"] = [("ground", 0.6, 0.8, 1.0, o), ("ground", 0.6, 0.0, 0.999999, o), ("ground", 0.6, 3.14159, 0.35, o),
                 ("tangent", 25_000.0, 0.0, o, 0.6), ("tangent", 1_000.0, 3.1, o, 0.6)]
    res = _run(c, sk.MultipleScatterSource.DiscreteOrdinates, sk.SingleScatterSource.Exact)
    ora = oracle.limb_radiance(**c, ms_do=True, ss_exact=True, exact_tangent=True)
    assert np.all(np.isfinite(res["radiance"]))
    np.testing.assert_allclose(res["radiance"][:, :, 0], ora["radiance"], rtol=1e-9)


def test_limb_plan_of_a_vertical_ray_does_not_index_out_of_range():
    """CPU part of the above: the host plan (and the oracle) for the degenerate SZA grid."""
    c = limb_case(nstr=4, nlayers=10, nwavel=2, nrays=2, seed=1)
    c["rays"] = [("ground", 0.6, 0.8, 1.0, 200_000.0)]
    geo, view = _geometry_and_view(c)
    ext = np.ascontiguousarray(c["ext"][:, 0])
    out = np.zeros((1, 6))
    npts = C.c_int(0)
    sza = np.zeros(2)
    _lib.check(_lib.lib().sk_b200_limb_plan_check(geo._geometry, view._viewing_geometry, 4, 2, _lib.dptr(ext), _lib.dptr(out),
                                                  C.byref(npts), _lib.dptr(sza)), "limb_plan_check")
    ora = oracle.limb_radiance(**c, ms_do=True, ss_exact=True, exact_tangent=True)
    np.testing.assert_allclose(out[0, 0], ora["los_optical_depth"][0, 0], rtol=1e-12)
    assert np.all(np.isfinite(ora["radiance"]))
