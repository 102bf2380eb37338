"""Pins the limb oracle (oracle/limb_oracle.hpp: spherical ray tracing, line-of-sight integration, interpolated DO
multiple-scatter source) to the reference's own golden numbers: tests/engine/test_1d_solver_regression.py:9-239
("discrete_ordinates" case: spherical geometry, 8 streams, num_sza = 2, two ground-viewing and two limb rays, three
wavelengths; radiances rtol 5e-7, line-of-sight optical depths rtol 5e-7).  Inputs restated from `_setup_1d`."""
import numpy as np
import pytest

from oracle import oracle

# expected_radiance / expected_optical_depth of the reference test (test_1d_solver_regression.py:118-139, 213-232)
GOLDEN_RADIANCE = np.array([
    [0.007325034080894165, 0.003911222331552876, 0.015579706096361565, 0.0038272940915240537],
    [0.015246498141181428, 0.009498372748059328, 0.024575183549304623, 0.007446741685194283],
    [0.023780265249057513, 0.016356191357612092, 0.031606480325730144, 0.01139900426172521],
])
GOLDEN_OPTICAL_DEPTH = np.array([
    [0.4046672641890289, 0.16756248699288395, 1.8429222725097874, 0.2573659860040063],
    [0.581709192271729, 0.2408710750522707, 2.64920076673282, 0.3699636048807591],
    [0.7587511203544293, 0.31417966311165746, 3.455479260955851, 0.4825612237575118],
])


def reference_case(num_wavelengths=3):
    """`_setup_1d("discrete_ordinates", 1, ...)` of the reference test (:9-104)."""
    alt = np.linspace(0.0, 60_000.0, 25)
    cos_sza = 0.42
    rays = [("ground", cos_sza, -0.7, 0.32, 200_000.0), ("ground", cos_sza, 0.4, 0.78, 200_000.0),
            ("tangent", 12_345.0, -0.35, 200_000.0, cos_sza), ("tangent", 27_123.0, 0.65, 200_000.0, cos_sza)]
    altitude_factor = np.exp(-alt / 7_500.0)[:, None]
    spectral = np.linspace(0.72, 1.35, num_wavelengths)[None, :]
    ext = (2.4e-5 * altitude_factor + 1.0e-9) * spectral
    ssa = 0.91 + 0.025 * np.exp(-alt / 18_000.0)[:, None] - 0.01 * np.linspace(0.0, 1.0, num_wavelengths)[None, :]
    leg = np.zeros((16, alt.size, num_wavelengths))
    leg[0], leg[1], leg[2] = 1.0, 0.08, 0.5
    albedo = np.linspace(0.08, 0.31, num_wavelengths)
    return dict(nstr=8, alt=alt, interp=1, cos_sza=cos_sza, saa=0.35, earth_radius=6_372_000.0, rays=rays, num_sza=2,
                ssa=ssa, ext=ext, leg=leg, albedo=albedo)


def test_limb_oracle_reproduces_reference_golden_optical_depths():
    out = oracle.limb_radiance(**reference_case(), ms_do=False, ss_exact=False)
    np.testing.assert_allclose(out["los_optical_depth"], GOLDEN_OPTICAL_DEPTH, rtol=5e-7, atol=1e-13)


def test_limb_oracle_reproduces_reference_golden_radiances():
    # single_scatter_source = DiscreteOrdinates in spherical geometry adds no line-of-sight single-scatter term
    # (cpp/lib/engine/engine.cpp:210-232): the golden radiance is the interpolated DO source alone
    out = oracle.limb_radiance(**reference_case(), ms_do=True, ss_exact=False)
    np.testing.assert_allclose(out["radiance"], GOLDEN_RADIANCE, rtol=5e-7, atol=2e-13)
    print("max rel diff vs golden:", np.max(np.abs(out["radiance"] / GOLDEN_RADIANCE - 1)))


def test_limb_oracle_optical_depth_equals_numerical_integration():
    """Independent check of the ray tracer + optical-depth quadrature (raytracing.h:478-560): the limb optical depth of
    a piecewise-linear extinction profile against adaptive numerical integration along the straight ray, 1e-10.  (The
    reference's golden limb optical depths sit 2.1e-7 away from this integral - inside its own 5e-7 tolerance.)"""
    from scipy.integrate import quad

    c = reference_case()
    alt, R, ext = c["alt"], c["earth_radius"], c["ext"][:, 0]
    out = oracle.limb_radiance(**c, ms_do=False, exact_tangent=True)["los_optical_depth"][0]
    for ray, ht in ((2, 12_345.0), (3, 27_123.0)):
        rt = R + ht
        rs = [rt] + [R + a for a in alt if R + a > rt]
        tot = 0.0
        for r0, r1 in zip(rs[:-1], rs[1:]):
            s0, s1 = np.sqrt(max(r0 * r0 - rt * rt, 0.0)), np.sqrt(r1 * r1 - rt * rt)
            tot += quad(lambda s: np.interp(np.sqrt(rt * rt + s * s) - R, alt, ext), s0, s1, epsabs=0, epsrel=1e-13)[0]
        assert abs(out[ray] / (2 * tot) - 1) < 1e-10


def test_limb_oracle_exact_single_scatter_thin_limit():
    """Exact single-scatter source (singlescattersource.cpp:949-1167) in the optically thin limit: the limb radiance tends
    to  sum over the path of  k omega P(Theta) / (4 pi) ds  with unattenuated sunlight."""
    c = reference_case(1)
    c["ext"] = c["ext"] * 1e-6
    out = oracle.limb_radiance(**c, ms_do=False, ss_exact=True)
    geo = oracle.limb_geometry(alt=c["alt"], interp=1, cos_sza=c["cos_sza"], saa=c["saa"], rays=c["rays"])
    R, alt = c["earth_radius"], c["alt"]
    for ray in (2, 3):
        n = geo["nlayers"][ray]
        lay = geo["layers"][ray, :n]
        # k * omega at both ends of every layer, path-length weighted by the two quadrature coefficients
        kw = lambda r: np.interp(r - R, alt, c["ext"][:, 0]) * np.interp(r - R, alt, c["ssa"][:, 0])  # noqa: E731
        total = np.sum(lay[:, 1] * kw(lay[:, 7]) + lay[:, 2] * kw(lay[:, 8]))
        # phase function 1 + 0.08 P1 + 0.5 P2 at the ray's scattering angle
        x = geo["cos_scatter"][ray]
        phase = 1 + 0.08 * x + 0.5 * 0.5 * (3 * x * x - 1)
        assert abs(out["radiance"][0, ray] / (phase / (4 * np.pi) * total) - 1) < 2e-5   # optical depths ~1e-6: T = 1 - O(1e-5)


def test_reference_tangent_layer_arithmetic_carries_rounding_noise():
    """Documents why limb parity against the reference cannot be tighter than its own 5e-7: the reference measures a
    tangent layer from sqrt(max(r_tan^2 - rt^2, 0)) with r_tan one rounding away from rt (spherical_shell.cpp:176-184),
    which shortens both tangent layers by ~0.1 m for about half of all rays.  On 100 tangent altitudes the restated
    arithmetic deviates from the exact-tangent variant by 1e-8 .. 3e-7 on a sizeable fraction of the rays and by
    nothing on the others; the exact-tangent variant equals numerical integration (previous test)."""
    alt = np.linspace(0.0, 100e3, 101)
    ext = (7e-5 * np.exp(-alt / 7400.0))[:, None]
    kw = dict(nstr=2, alt=alt, interp=1, cos_sza=0.6, rays=[("tangent", float(h), 0.3, 200e3, 0.6) for h in np.linspace(10e3, 60e3, 100)],
              num_sza=1, ms_do=False, ssa=np.full_like(ext, 0.9), ext=ext, leg=np.ones((1, alt.size, 1)), albedo=np.zeros(1))
    exact = oracle.limb_radiance(**kw, exact_tangent=True)["los_optical_depth"][0]
    ref = oracle.limb_radiance(**kw, exact_tangent=False)["los_optical_depth"][0]
    dev = np.abs(ref / exact - 1)
    assert dev.max() < 5e-7
    assert np.sum(dev > 1e-9) >= 10 and np.sum(dev < 1e-12) >= 10
