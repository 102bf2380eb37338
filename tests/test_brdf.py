"""Non-Lambertian surface BRDFs: the MODIS model (isotropic + Ross-thick + Li-sparse-R) and the snow
model of Kokhanovsky (cpp/include/sasktran2/atmosphere/surface.h:140-362) through the discrete-ordinates solve.

The reference holds no numbers for it (tests/constituent/test_modis.py only runs it and checks finite differences), so
the oracle's restatement of the models and of SurfaceStorage::compute_expansion (sktran_do_surface.h:49-91) is pinned by
identities: an isotropic-only MODIS surface is the Lambertian one, the Fourier series reproduces the model, the kernels
are reciprocal.  The CUDA path - which expands the kernels once on the host and combines them per wavelength on the
device - is then compared with the oracle's per-wavelength quadrature."""
import os

import numpy as np
import pytest

import sasktran2_b200 as sk
from oracle import oracle
from sasktran2_b200 import _lib, scenarios


def _oracle_kw(sc):
    return dict(nstr=sc.nstr, alt=sc.altitudes, interp=sc.interp, geotype=sc.geotype, cos_sza=sc.cos_sza,
                los_cos_vza=sc.los_cos_vza, los_rel_az=sc.los_rel_az, ssa=sc.ssa, ext=sc.total_extinction, leg=sc.leg_coeff,
                albedo=sc.albedo)


def _case(nstr=8, nwavel=4, nlos=3):
    sc = scenarios.small_wf_case(nstr=nstr, nlayers=10, nwavel=nwavel, nlos=nlos)
    sc.mappings = {}
    args = np.zeros((3, nwavel))
    args[0] = np.linspace(0.1, 0.4, nwavel)
    args[1] = np.linspace(0.02, 0.08, nwavel)
    args[2] = np.linspace(0.05, 0.01, nwavel)
    return sc, args


def test_oracle_modis_isotropic_only_is_lambertian():
    sc, args = _case()
    args[1:] = 0.0
    sc.albedo = args[0].copy()
    a = oracle.do_radiance(**_oracle_kw(sc))["radiance"]
    b = oracle.do_radiance(**_oracle_kw(sc), brdf_kind=2, brdf_args=args)["radiance"]
    np.testing.assert_allclose(b, a, rtol=1e-12)


def test_oracle_brdf_expansion_reproduces_the_model_and_is_reciprocal():
    args = [0.2, 0.05, 0.03]
    for mu_i, mu_o, phi in ((0.6, 0.4, 0.7), (0.9, 0.3, 2.5), (0.5, 0.5, 0.1)):
        series = sum(oracle.brdf_expansion(m, 2, args, mu_o, mu_i) * np.cos(m * phi) for m in range(200))
        assert abs(series - np.pi * oracle.brdf_value(2, args, mu_i, mu_o, phi)) < 2e-6
        for m in (0, 1, 5):
            assert abs(oracle.brdf_expansion(m, 2, args, mu_o, mu_i) - oracle.brdf_expansion(m, 2, args, mu_i, mu_o)) < 1e-13
    # the snow model at normal incidence and the Lambertian special case
    assert abs(oracle.brdf_expansion(0, 0, [0.3], 0.5, 0.7) - 0.3) < 1e-15 and oracle.brdf_expansion(2, 0, [0.3], 0.5, 0.7) == 0.0
    assert oracle.brdf_value(1, [1e-6], 0.6, 0.7, 1.0) > 0


def _run(sc, args=None, calc_derivs=False):
    cfg = sk.Config()
    cfg.num_streams = sc.nstr
    cfg.multiple_scatter_source = sk.MultipleScatterSource.DiscreteOrdinates
    cfg.single_scatter_source = sk.SingleScatterSource.DiscreteOrdinates
    geo = sk.Geometry1D(sc.cos_sza, 0.0, sc.earth_radius, sc.altitudes, sk.InterpolationMethod(sc.interp), sk.GeometryType(sc.geotype))
    view = sk.ViewingGeometry()
    for cz, az in zip(sc.los_cos_vza, sc.los_rel_az):
        view.add_ray(sk.GroundViewingSolar(sc.cos_sza, float(az), float(cz), sc.observer_altitude))
    eng = sk.Engine(cfg, geo, view)
    atm = sk.Atmosphere.from_scenario(sc, geo, cfg, calculate_derivatives=calc_derivs)
    if args is not None:
        atm.surface.use_modis(args[0], args[1], args[2])
    if calc_derivs:
        atm.surface.enable_albedo_derivative("wf_albedo")
    return eng.calculate_radiance(atm)


@pytest.mark.gpu
@pytest.mark.parametrize("nstr,generic", [(4, False), (8, False), (16, False), (8, True), (2, False)])
def test_cuda_modis_brdf_vs_oracle(nstr, generic):
    sc, args = _case(nstr=nstr)
    if generic:
        os.environ["SK_B200_GENERIC"] = "1"
    try:
        rad = _run(sc, args)["radiance"][:, :, 0]
    finally:
        os.environ.pop("SK_B200_GENERIC", None)
    ora = oracle.do_radiance(**_oracle_kw(sc), brdf_kind=2, brdf_args=args)["radiance"]
    err = np.max(np.abs(rad / ora - 1))
    print(f"MODIS nstr={nstr} generic={generic}: max rel diff vs oracle {err:.2e}")
    assert err < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("nstr", [4, 8, 16])
def test_cuda_snow_brdf_vs_oracle(nstr):
    """Snow BRDF of Kokhanovsky (not linear in its argument): device quadrature per wavelength against the oracle's."""
    sc, _ = _case(nstr=nstr)
    arg = np.linspace(2e-7, 5e-6, sc.nwavel)[None, :]

    def run():
        cfg = sk.Config()
        cfg.num_streams = sc.nstr
        cfg.multiple_scatter_source = sk.MultipleScatterSource.DiscreteOrdinates
        cfg.single_scatter_source = sk.SingleScatterSource.DiscreteOrdinates
        geo = sk.Geometry1D(sc.cos_sza, 0.0, sc.earth_radius, sc.altitudes, sk.InterpolationMethod(sc.interp), sk.GeometryType(sc.geotype))
        view = sk.ViewingGeometry()
        for cz, az in zip(sc.los_cos_vza, sc.los_rel_az):
            view.add_ray(sk.GroundViewingSolar(sc.cos_sza, float(az), float(cz), sc.observer_altitude))
        eng = sk.Engine(cfg, geo, view)
        atm = sk.Atmosphere.from_scenario(sc, geo, cfg, calculate_derivatives=False)
        atm.surface.use_snow_kokhanovsky(arg[0])
        return eng.calculate_radiance(atm)["radiance"][:, :, 0]

    rad = run()
    ora = oracle.do_radiance(**_oracle_kw(sc), brdf_kind=1, brdf_args=arg)["radiance"]
    err = np.max(np.abs(rad / ora - 1))
    print(f"snow nstr={nstr}: max rel diff vs oracle {err:.2e}")
    assert err < 1e-9


@pytest.mark.gpu
def test_cuda_modis_isotropic_only_equals_lambertian_path():
    sc, args = _case(nstr=8)
    args[1:] = 0.0
    sc.albedo = args[0].copy()
    lam = _run(sc)["radiance"]
    mod = _run(sc, args)["radiance"]
    np.testing.assert_allclose(mod, lam, rtol=1e-11)


@pytest.mark.gpu
def test_cuda_non_lambertian_refusals():
    sc, args = _case(nstr=4)
    sc.mappings = scenarios.small_wf_case(nstr=4, nlayers=10, nwavel=4, nlos=3).mappings
    with pytest.raises(_lib.SasktranError):     # weighting functions with a kernel-based BRDF
        _run(sc, args, calc_derivs=True)
