"""Thermal emission solved by the discrete-ordinates path (config.emission_source = DiscreteOrdinates; SURVEY row a2 /
a7: OpticalLayerArray's b0 / b1, RTESolver::solveParticularGreenThermal, the thermal terms of
OpticalLayer::integrate_source, the surface emission of the ground boundary).

Pins: the two numbers of the reference's own test (tests/engine/thermal_emissions/test_disort.py:8-112, DISORT test
case 7a with modifications, rtol 1e-6 there), the analytic solution of the non-scattering limit, and the invariance of
an exponential emission profile under splitting its layer (the b1 != 0 terms).  Then the kernel bodies (host emulation)
and the CUDA path against the oracle."""
import os

import numpy as np
import pytest

import sasktran2_b200 as sk
from oracle import oracle
from sasktran2_b200 import scenarios

from .test_host_emulation import emul  # noqa: F401  (fixture)

B_REF = 1.09657540e-05


def _disort_7a(od, ssa0=0.95, g=0.75):
    alt = np.array([0.0, 1000.0])
    ext = np.full((2, 1), od / 1000.0)
    ssa = np.full((2, 1), ssa0)
    leg = np.zeros((17, 2, 1))
    for l in range(17):
        leg[l] = g ** l * (2 * l + 1)
    return alt, ext, ssa, leg


@pytest.mark.parametrize("od,surface,expected", [(100.0, False, 7.93075833e-06), (1.0, True, 1.02396134e-05)])
def test_oracle_reproduces_the_reference_thermal_numbers(od, surface, expected):
    alt, ext, ssa, leg = _disort_7a(od)
    sc = oracle.apply_delta_m_scaling(16, ssa, ext, leg)
    r = oracle.do_radiance(nstr=16, alt=alt, interp=1, geotype=0, cos_sza=0.6, los_cos_vza=[1.0], los_rel_az=[0.0],
                           ssa=sc["ssa"], ext=sc["ext"], leg=sc["leg"], f=sc["f"], solar=np.zeros(1), albedo=0.0,
                           emission=np.full((2, 1), B_REF), surface_emission=(B_REF if surface else None))
    np.testing.assert_allclose(r["radiance"][0, 0], expected, rtol=1e-6)   # the reference's own tolerance


def test_oracle_non_scattering_limit_is_the_schwarzschild_solution():
    """omega -> 0 (the reference's dither floor 1e-9... here 1e-12 of scattering): I = B (1 - e^(-tau/mu)) + B_s e^(-tau/mu)
    for an isothermal slab, and the integral of b0 e^(-b1 x) e^(-x/mu) / mu for an exponential profile."""
    alt = np.linspace(0.0, 3000.0, 4)
    nloc = alt.size
    ext = np.full((nloc, 1), 4e-4)
    ssa = np.full((nloc, 1), 1e-12)
    leg = np.zeros((4, nloc, 1))
    leg[0] = 1.0
    mus = np.array([1.0, 0.7, 0.4])
    kw = dict(nstr=4, alt=alt, interp=1, geotype=0, cos_sza=0.6, los_cos_vza=mus, los_rel_az=np.zeros(3), ssa=ssa, ext=ext,
              leg=leg, solar=np.zeros(1), albedo=0.0)
    tau = 4e-4 * 3000.0
    r = oracle.do_radiance(**kw, emission=np.full((nloc, 1), 2.0), surface_emission=5.0)["radiance"][0]
    np.testing.assert_allclose(r, 2.0 * (1 - np.exp(-tau / mus)) + 5.0 * np.exp(-tau / mus), rtol=1e-8)
    # exponential in optical depth from the top: B(x) = B_top exp(-b x); grid values at x = tau, 2 tau / 3, tau / 3, 0
    b = 0.9
    x_grid = tau * (1 - alt / 3000.0)
    em = (3.0 * np.exp(-b * x_grid))[:, None]
    r = oracle.do_radiance(**kw, emission=em)["radiance"][0]
    exact = 3.0 / (1 + mus * b) * (1 - np.exp(-tau * (b + 1 / mus)))
    np.testing.assert_allclose(r, exact, rtol=1e-8)


def test_oracle_exponential_profile_is_invariant_under_layer_splitting():
    """A layer whose emission falls off exponentially in optical depth is the same medium as its two halves with the
    geometric-mean emission on the new level: exercises every b1 != 0 branch with scattering on."""
    def run(alt, em):
        nloc = alt.size
        ext = np.full((nloc, 1), 8e-4)
        ssa = np.full((nloc, 1), 0.8)
        leg = np.zeros((8, nloc, 1))
        for l in range(8):
            leg[l] = 0.6 ** l * (2 * l + 1)
        return oracle.do_radiance(nstr=8, alt=alt, interp=1, geotype=0, cos_sza=0.5, los_cos_vza=[1.0, 0.6, 0.3],
                                  los_rel_az=[0.0, 1.0, 2.0], ssa=ssa, ext=ext, leg=leg, solar=np.ones(1), albedo=0.3,
                                  emission=em[:, None], surface_emission=0.7)["radiance"][0]
    one = run(np.array([0.0, 1000.0]), np.array([1.0, 4.0]))
    two = run(np.array([0.0, 500.0, 1000.0]), np.array([1.0, 2.0, 4.0]))
    np.testing.assert_allclose(two, one, rtol=1e-10)


def _scenario(nstr, nwavel=5, nlos=4, nlayers=9, geotype=1):
    sc = scenarios.small_wf_case(nstr=nstr, nlayers=nlayers, nwavel=nwavel, nlos=nlos, geotype=geotype)
    sc.mappings = {}
    nloc = sc.altitudes.size
    # a temperature-like profile: emission varying by a factor of ~4 over the column and across the wavelengths,
    # one level repeated (b1 = 0 there) and solar scattering left on
    z = np.linspace(0.0, 1.0, nloc)
    em = (0.02 + 0.06 * np.exp(-3.0 * z))[:, None] * np.linspace(1.0, 2.0, nwavel)[None, :]
    em[3] = em[2]
    return sc, np.asfortranarray(em), np.linspace(0.05, 0.09, nwavel)


def _oracle_kw(sc):
    return dict(nstr=sc.nstr, alt=sc.altitudes, interp=sc.interp, geotype=sc.geotype, cos_sza=sc.cos_sza,
                los_cos_vza=sc.los_cos_vza, los_rel_az=sc.los_rel_az, ssa=sc.ssa, ext=sc.total_extinction, leg=sc.leg_coeff,
                albedo=sc.albedo)


@pytest.mark.parametrize("nstr", [2, 4, 8, 16])
def test_kernel_bodies_with_thermal_emission_match_the_oracle(emul, nstr):  # noqa: F811
    sc, em, se = _scenario(nstr)
    ref = oracle.do_radiance(**_oracle_kw(sc), emission=em, surface_emission=se)["radiance"]
    no_emission = oracle.do_radiance(**_oracle_kw(sc))["radiance"]
    assert np.all(ref > no_emission * 1.05)   # the thermal part is not a rounding-level contribution
    rad, _ = emul(**_oracle_kw(sc), emission=em, surface_emission=se)
    np.testing.assert_allclose(rad, ref, rtol=1e-10)
    # surface emission alone (emission_source = NoSource but a non-zero Surface.emission)
    rad, _ = emul(**_oracle_kw(sc), surface_emission=se)
    np.testing.assert_allclose(rad, oracle.do_radiance(**_oracle_kw(sc), surface_emission=se)["radiance"], rtol=1e-10)


# ---------------------------------------------------------------------------------------------------------------------
# CUDA
# ---------------------------------------------------------------------------------------------------------------------
def _run_cuda(sc, em=None, se=None, emission_source=sk.EmissionSource.DiscreteOrdinates, calc_derivs=False):
    cfg = sk.Config()
    cfg.num_streams = sc.nstr
    cfg.multiple_scatter_source = sk.MultipleScatterSource.DiscreteOrdinates
    cfg.single_scatter_source = sk.SingleScatterSource.DiscreteOrdinates
    cfg.emission_source = emission_source
    geo = sk.Geometry1D(sc.cos_sza, 0.0, sc.earth_radius, sc.altitudes, sk.InterpolationMethod(sc.interp), sk.GeometryType(sc.geotype))
    view = sk.ViewingGeometry()
    for cz, az in zip(sc.los_cos_vza, sc.los_rel_az):
        view.add_ray(sk.GroundViewingSolar(sc.cos_sza, float(az), float(cz), sc.observer_altitude))
    eng = sk.Engine(cfg, geo, view)
    atm = sk.Atmosphere.from_scenario(sc, geo, cfg, calculate_derivatives=calc_derivs)
    if em is not None:
        atm.storage.emission_source[:] = em
    if se is not None:
        atm.surface.emission[:] = se
    return eng.calculate_radiance(atm)


@pytest.mark.gpu
@pytest.mark.parametrize("od,surface,expected", [(100.0, False, 7.93075833e-06), (1.0, True, 1.02396134e-05)])
def test_cuda_reproduces_the_reference_thermal_numbers(od, surface, expected):
    """The reference's own test, call for call (tests/engine/thermal_emissions/test_disort.py)."""
    alt, ext, ssa, leg = _disort_7a(od)
    cfg = sk.Config()
    cfg.multiple_scatter_source = sk.MultipleScatterSource.DiscreteOrdinates
    cfg.single_scatter_source = sk.SingleScatterSource.DiscreteOrdinates
    cfg.emission_source = sk.EmissionSource.DiscreteOrdinates
    cfg.num_streams = 16
    cfg.num_singlescatter_moments = 17
    cfg.delta_m_scaling = True   # the reference test calls apply_delta_m_scaling(num_streams) by hand
    geo = sk.Geometry1D(0.6, 0.0, 6372000.0, alt, sk.InterpolationMethod.LinearInterpolation, sk.GeometryType.PlaneParallel)
    view = sk.ViewingGeometry()
    view.add_ray(sk.GroundViewingSolar(0.6, 0.0, 1.0, 200000.0))
    atm = sk.Atmosphere(geo, cfg, numwavel=1, calculate_derivatives=False, num_legendre=17)
    atm.storage.total_extinction[:] = ext
    atm.storage.ssa[:] = ssa
    atm.storage.solar_irradiance[:] = 0.0
    atm.storage.emission_source[:] = B_REF
    if surface:
        atm.surface.emission[:] = B_REF
    atm.storage.leg_coeff[:] = leg
    rad = sk.Engine(cfg, geo, view).calculate_radiance(atm)["radiance"]
    np.testing.assert_allclose(rad[0, 0, 0], expected, rtol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("nstr,generic,geotype", [(4, False, 1), (8, False, 0), (16, False, 1), (16, True, 1), (2, False, 1),
                                                  (32, False, 0)])
def test_cuda_thermal_emission_vs_oracle(nstr, generic, geotype):
    sc, em, se = _scenario(nstr, geotype=geotype)
    if generic:
        os.environ["SK_B200_GENERIC"] = "1"
    try:
        rad = _run_cuda(sc, em, se)["radiance"][:, :, 0]
        only_surface = _run_cuda(sc, None, se, emission_source=sk.EmissionSource.NoSource)["radiance"][:, :, 0]
    finally:
        os.environ.pop("SK_B200_GENERIC", None)
    ref = oracle.do_radiance(**_oracle_kw(sc), emission=em, surface_emission=se)["radiance"]
    err = np.max(np.abs(rad / ref - 1))
    print(f"thermal nstr={nstr} generic={generic}: max rel diff vs oracle {err:.2e}")
    assert err < 1e-9
    ref = oracle.do_radiance(**_oracle_kw(sc), surface_emission=se)["radiance"]
    assert np.max(np.abs(only_surface / ref - 1)) < 1e-9


@pytest.mark.gpu
def test_cuda_thermal_emission_chunked_equals_unchunked():
    sc, em, se = _scenario(8, nwavel=37)
    a = _run_cuda(sc, em, se)["radiance"]
    os.environ["SK_B200_WORKSPACE_GB"] = "0.002"
    try:
        b = _run_cuda(sc, em, se)["radiance"]
    finally:
        os.environ.pop("SK_B200_WORKSPACE_GB", None)
    assert np.array_equal(a, b)


@pytest.mark.gpu
def test_cuda_refuses_weighting_functions_with_thermal_emission():
    sc = scenarios.small_wf_case(nstr=8, nlayers=6, nwavel=3, nlos=2)
    em = np.full((sc.altitudes.size, sc.nwavel), 0.01)
    with pytest.raises(sk.SasktranError, match="thermal emission"):
        cfg = sk.Config()
        cfg.num_streams = sc.nstr
        cfg.multiple_scatter_source = sk.MultipleScatterSource.DiscreteOrdinates
        cfg.single_scatter_source = sk.SingleScatterSource.DiscreteOrdinates
        cfg.emission_source = sk.EmissionSource.DiscreteOrdinates
        geo = sk.Geometry1D(sc.cos_sza, 0.0, sc.earth_radius, sc.altitudes, sk.InterpolationMethod(sc.interp), sk.GeometryType(sc.geotype))
        view = sk.ViewingGeometry()
        for cz, az in zip(sc.los_cos_vza, sc.los_rel_az):
            view.add_ray(sk.GroundViewingSolar(sc.cos_sza, float(az), float(cz), sc.observer_altitude))
        eng = sk.Engine(cfg, geo, view)
        atm = sk.Atmosphere.from_scenario(sc, geo, cfg, calculate_derivatives=True)
        atm.storage.emission_source[:] = em
        eng.calculate_radiance(atm)


@pytest.mark.gpu
def test_cuda_emission_do_requires_do_scattering_sources():
    """Config::validate_config (cpp/lib/config/config.cpp:127-141; tests/input_validation/test_emission_validation.py)."""
    sc = scenarios.small_wf_case(nstr=4, nlayers=5, nwavel=2, nlos=1)
    geo = sk.Geometry1D(sc.cos_sza, 0.0, sc.earth_radius, sc.altitudes, sk.InterpolationMethod(sc.interp), sk.GeometryType(sc.geotype))
    view = sk.ViewingGeometry()
    view.add_ray(sk.GroundViewingSolar(sc.cos_sza, 0.0, 1.0, sc.observer_altitude))
    cfg = sk.Config()
    cfg.num_streams = 4
    cfg.multiple_scatter_source = sk.MultipleScatterSource.DiscreteOrdinates
    cfg.single_scatter_source = sk.SingleScatterSource.NoSource
    cfg.emission_source = sk.EmissionSource.DiscreteOrdinates
    with pytest.raises(sk.SasktranError, match="requires single_scatter_source"):
        sk.Engine(cfg, geo, view)
