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
    """The snow model is not linear in its argument: a weighting function w.r.t. it is refused (MODIS weights are solved)."""
    sc, _ = _case(nstr=4)
    sc.mappings = scenarios.small_wf_case(nstr=4, nlayers=10, nwavel=4, nlos=3).mappings
    cfg = sk.Config()
    cfg.num_streams = sc.nstr
    cfg.multiple_scatter_source = sk.MultipleScatterSource.DiscreteOrdinates
    cfg.single_scatter_source = sk.SingleScatterSource.DiscreteOrdinates
    geo = sk.Geometry1D(sc.cos_sza, 0.0, sc.earth_radius, sc.altitudes, sk.InterpolationMethod(sc.interp), sk.GeometryType(sc.geotype))
    view = sk.ViewingGeometry()
    for cz, az in zip(sc.los_cos_vza, sc.los_rel_az):
        view.add_ray(sk.GroundViewingSolar(sc.cos_sza, float(az), float(cz), sc.observer_altitude))
    eng = sk.Engine(cfg, geo, view)
    atm = sk.Atmosphere.from_scenario(sc, geo, cfg, calculate_derivatives=True)
    atm.surface.use_snow_kokhanovsky(np.linspace(2e-7, 5e-6, sc.nwavel))
    atm.surface.enable_brdf_argument_derivative("wf_snow", 0, num_args=1)
    with pytest.raises(_lib.SasktranError, match="snow BRDF"):
        eng.calculate_radiance(atm)


# ---------------------------------------------------------------------------------------------------------------------
# weighting functions of the atmosphere above a non-Lambertian surface
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", [2, 1], ids=["modis", "snow"])
def test_oracle_atmospheric_wf_above_a_brdf_surface_match_finite_differences(kind):
    """The BRDF does not depend on the atmosphere, so the forward-mode lanes of the oracle go through the general ground
    rows unchanged; checked against central differences of its own radiances (the reference's criterion for weighting
    functions, src/sasktran2/test_util/wf.py:9-80)."""
    from .test_oracle_wf import species_atmosphere

    nstr, nlayers, nwavel = 8, 8, 2
    z, build, k_aer0, k_abs0, w_aer, b_aer = species_atmosphere(nstr, nlayers, nwavel)
    nloc = z.size
    cz, az = np.array([0.9, 0.5]), np.array([0.3, 2.1])
    args = (np.array([[0.25, 0.3], [0.06, 0.05], [0.03, 0.02]]) if kind == 2 else np.array([[2e-6, 4e-6]]))
    common = dict(nstr=nstr, alt=z, interp=2, geotype=1, cos_sza=0.55, los_cos_vza=cz, los_rel_az=az, albedo=0.0,
                  brdf_kind=kind, brdf_args=args)
    k, ssa, leg, ks = build(k_aer0, k_abs0)
    d_leg = (b_aer[:, None, None] - leg)[..., None]
    base = oracle.do_radiance(**common, ssa=ssa, ext=k, leg=leg, d_leg=d_leg, calc_derivs=True)
    maps = {"abs": dict(d_extinction=np.ones_like(k), d_ssa=-ssa / k),
            "aer": dict(d_extinction=np.ones_like(k), d_ssa=(w_aer - ssa) / k, scat_factor=w_aer / ks, scat_index=0)}
    wf = oracle.apply_mappings(base["native"], maps, nloc, 1)

    def rad(k_aer, k_abs):
        kk, ss, ll, _ = build(k_aer, k_abs)
        return oracle.do_radiance(**common, ssa=ss, ext=kk, leg=ll)["radiance"]

    for q in (0, 3, nlayers - 1):
        for name, which in (("abs", 1), ("aer", 0)):
            fd = np.zeros((nwavel, cz.size))
            for w in range(nwavel):
                pert = [k_aer0.copy(), k_abs0.copy()]
                h = 1e-4 * k[q, w]
                pert[which][q, w] += h
                up = rad(*pert)[w]
                pert[which][q, w] -= 2 * h
                dn = rad(*pert)[w]
                fd[w] = (up - dn) / (2 * h)
            scale = np.abs(wf[name]).max(axis=0)
            assert np.abs(wf[name][q] - fd).max() / scale.max() < 5e-6, (name, q, wf[name][q], fd)   # FD truncation


def _oracle_wf_brdf(sc, kind, args):
    from . import wf_checks

    names = wf_checks.scat_names(sc)
    d_leg = np.stack([sc.mappings[n]["d_legendre"] for n in names], axis=-1) if names else None
    ora = oracle.do_radiance(**wf_checks.oracle_inputs(sc), d_leg=d_leg, calc_derivs=True, stable=True, reverse=False,
                             brdf_kind=kind, brdf_args=args)
    maps = {n: dict(d_ssa=mp["d_ssa"], d_extinction=mp["d_extinction"], scat_factor=mp.get("scat_factor"),
                    scat_index=names.index(n) if n in names else -1, interpolator=mp.get("interpolator"))
            for n, mp in sc.mappings.items()}
    return ora["radiance"], oracle.apply_mappings(ora["native"], maps, sc.nloc, len(names))


def _run_wf(sc, kind, args):
    cfg = sk.Config()
    cfg.num_streams = sc.nstr
    cfg.multiple_scatter_source = sk.MultipleScatterSource.DiscreteOrdinates
    cfg.single_scatter_source = sk.SingleScatterSource.DiscreteOrdinates
    geo = sk.Geometry1D(sc.cos_sza, 0.0, sc.earth_radius, sc.altitudes, sk.InterpolationMethod(sc.interp), sk.GeometryType(sc.geotype))
    view = sk.ViewingGeometry()
    for cz, az in zip(sc.los_cos_vza, sc.los_rel_az):
        view.add_ray(sk.GroundViewingSolar(sc.cos_sza, float(az), float(cz), sc.observer_altitude))
    eng = sk.Engine(cfg, geo, view)
    atm = sk.Atmosphere.from_scenario(sc, geo, cfg, calculate_derivatives=True)
    if kind == 2:
        atm.surface.use_modis(args[0], args[1], args[2])
    else:
        atm.surface.use_snow_kokhanovsky(args[0])
    return eng.calculate_radiance(atm)


@pytest.mark.gpu
@pytest.mark.parametrize("nstr,kind,generic,nlos", [(8, 2, False, 4), (16, 2, False, 5), (4, 2, False, 9), (16, 1, False, 3),
                                                    (16, 2, True, 4), (16, 2, False, 24)])
def test_cuda_atmospheric_wf_above_a_brdf_surface_vs_oracle(nstr, kind, generic, nlos):
    """Weighting functions of the atmosphere (absorbers, a scatterer group) above a MODIS / snow surface: the layer on the
    ground goes through k_wf_layer with the reflection rows of k_surface_general (every order reflects), the rest through
    the register-resident kernel; against the oracle's forward-mode lanes, 1e-7 of each column maximum."""
    sc = scenarios.small_wf_case(nstr=nstr, nlayers=9, nwavel=4, nlos=nlos)
    if kind == 2:
        args = np.zeros((3, sc.nwavel))
        args[0] = np.linspace(0.1, 0.4, sc.nwavel)
        args[1] = np.linspace(0.02, 0.08, sc.nwavel)
        args[2] = np.linspace(0.05, 0.01, sc.nwavel)
    else:
        args = np.linspace(2e-7, 5e-6, sc.nwavel)[None, :]
    if generic:
        os.environ["SK_B200_GENERIC"] = "1"
    try:
        res = _run_wf(sc, kind, args)
    finally:
        os.environ.pop("SK_B200_GENERIC", None)
    rad, wf = _oracle_wf_brdf(sc, kind, args)
    assert np.max(np.abs(res["radiance"][:, :, 0] / rad - 1)) < 1e-9
    worst = 0.0
    for name, ref in wf.items():
        got = res[name][..., 0]
        err = np.abs(got - ref) / np.abs(ref).max(axis=0, keepdims=True)
        tol = 1e-5 if "aerosol" in name else 1e-7
        worst = max(worst, float(err.max()) / tol)
        assert err.max() < tol, (name, float(err.max()))
    print(f"BRDF kind {kind} nstr={nstr} generic={generic} nlos={nlos}: worst WF error / tolerance {worst:.2e}")


# the product's kernel bodies on the host (tests/host_emul.cpp): MODIS surface, radiances and native derivatives
from .test_host_emulation import emul  # noqa: E402,F401  (fixture)


@pytest.mark.parametrize("nstr,nlos", [(4, 3), (8, 2), (2, 2), (16, 5)])
def test_kernel_bodies_modis_surface_and_wf_above_it_match_the_oracle(emul, nstr, nlos):  # noqa: F811
    from . import wf_checks

    sc = scenarios.small_wf_case(nstr=nstr, nlayers=7, nwavel=3, nlos=nlos)
    args = np.zeros((3, sc.nwavel))
    args[0] = np.linspace(0.1, 0.4, sc.nwavel)
    args[1] = np.linspace(0.02, 0.08, sc.nwavel)
    args[2] = np.linspace(0.05, 0.01, sc.nwavel)
    names = wf_checks.scat_names(sc)
    d_leg = np.stack([sc.mappings[n]["d_legendre"] for n in names], axis=-1)
    kw = wf_checks.oracle_inputs(sc)
    ora = oracle.do_radiance(**kw, d_leg=d_leg, calc_derivs=True, stable=True, reverse=False, brdf_kind=2, brdf_args=args)
    rad, _, native = emul(**kw, d_leg=d_leg, want_native=True, modis_args=args)
    np.testing.assert_allclose(rad, ora["radiance"], rtol=1e-10)
    nloc = sc.nloc
    ref = ora["native"][:, :, :-1]        # the last lane is the Lambertian albedo (meaningless here)
    got = native[:, :, :-1]
    for block in range(ref.shape[2] // nloc):
        r, g = ref[:, :, block * nloc:(block + 1) * nloc], got[:, :, block * nloc:(block + 1) * nloc]
        scale = np.abs(r).max(axis=2, keepdims=True)
        assert np.max(np.abs(g - r) / scale) < 1e-7, (block, float(np.max(np.abs(g - r) / scale)))


def _modis_args(nwavel):
    args = np.zeros((3, nwavel))
    args[0] = np.linspace(0.1, 0.4, nwavel)
    args[1] = np.linspace(0.02, 0.08, nwavel)
    args[2] = np.linspace(0.05, 0.01, nwavel)
    return args


def _oracle_fd_wrt_kernel_weights(sc, args, h=1e-5):
    """Central differences of the oracle's radiances w.r.t. the three MODIS kernel weights: [3, nwavel, nlos] (the
    reference's own criterion for weighting functions, src/sasktran2/test_util/wf.py:9-80)."""
    out = []
    for k in range(3):
        up, dn = args.copy(), args.copy()
        up[k] += h
        dn[k] -= h
        out.append((oracle.do_radiance(**_oracle_kw(sc), brdf_kind=2, brdf_args=up)["radiance"] -
                    oracle.do_radiance(**_oracle_kw(sc), brdf_kind=2, brdf_args=dn)["radiance"]) / (2 * h))
    return np.array(out)


@pytest.mark.parametrize("nstr,nlos", [(4, 3), (8, 2), (16, 4)])
def test_kernel_bodies_modis_weight_derivatives_match_finite_differences(emul, nstr, nlos):  # noqa: F811
    import ctypes

    from . import wf_checks
    from .test_host_emulation import ROOT

    sc = scenarios.small_wf_case(nstr=nstr, nlayers=7, nwavel=3, nlos=nlos)
    args = _modis_args(sc.nwavel)
    names = wf_checks.scat_names(sc)
    d_leg = np.stack([sc.mappings[n]["d_legendre"] for n in names], axis=-1)
    emul(**wf_checks.oracle_inputs(sc), d_leg=d_leg, want_native=True, modis_args=args)
    hl = ctypes.CDLL(str(ROOT / "tests" / "libhost_emul.so"))
    hl.emul_dump.restype = ctypes.c_longlong
    buf = np.zeros(sc.nwavel * nlos * 3)
    n = hl.emul_dump(b"wf_gndk", buf.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), ctypes.c_longlong(buf.size))
    assert n == buf.size
    got = buf.reshape(sc.nwavel, nlos, 3).transpose(2, 0, 1)
    fd = _oracle_fd_wrt_kernel_weights(sc, args)
    np.testing.assert_allclose(got, fd, rtol=2e-7, atol=1e-9 * np.abs(fd).max())


@pytest.mark.gpu
@pytest.mark.parametrize("nstr,generic,nlos", [(8, False, 4), (16, False, 5), (16, True, 3), (4, False, 24)])
def test_cuda_modis_weight_derivatives_vs_finite_differences(nstr, generic, nlos):
    """d radiance / d (weight of MODIS kernel k) through the surface derivative mappings (d_brdf[:, k] = 1), next to the
    atmospheric weighting functions of the same call; against central differences of the oracle's radiances."""
    sc = scenarios.small_wf_case(nstr=nstr, nlayers=9, nwavel=4, nlos=nlos)
    args = _modis_args(sc.nwavel)
    cfg = sk.Config()
    cfg.num_streams = sc.nstr
    cfg.multiple_scatter_source = sk.MultipleScatterSource.DiscreteOrdinates
    cfg.single_scatter_source = sk.SingleScatterSource.DiscreteOrdinates
    geo = sk.Geometry1D(sc.cos_sza, 0.0, sc.earth_radius, sc.altitudes, sk.InterpolationMethod(sc.interp), sk.GeometryType(sc.geotype))
    view = sk.ViewingGeometry()
    for cz, az in zip(sc.los_cos_vza, sc.los_rel_az):
        view.add_ray(sk.GroundViewingSolar(sc.cos_sza, float(az), float(cz), sc.observer_altitude))
    if generic:
        os.environ["SK_B200_GENERIC"] = "1"
    try:
        eng = sk.Engine(cfg, geo, view)
        atm = sk.Atmosphere.from_scenario(sc, geo, cfg, calculate_derivatives=True)
        atm.surface.use_modis(args[0], args[1], args[2])
        for k, name in enumerate(("wf_brdf_iso", "wf_brdf_vol", "wf_brdf_geo")):
            atm.surface.enable_brdf_argument_derivative(name, k)
        res = eng.calculate_radiance(atm)
    finally:
        os.environ.pop("SK_B200_GENERIC", None)
    fd = _oracle_fd_wrt_kernel_weights(sc, args)
    for k, name in enumerate(("wf_brdf_iso", "wf_brdf_vol", "wf_brdf_geo")):
        got = np.asarray(res[name]).reshape(sc.nwavel, nlos)
        np.testing.assert_allclose(got, fd[k], rtol=2e-7, atol=1e-9 * np.abs(fd).max())
    # the atmospheric weighting functions of the same call are unaffected
    _, wf = _oracle_wf_brdf(sc, 2, args)
    err = np.abs(res["wf_o3_vmr"][..., 0] - wf["wf_o3_vmr"]) / np.abs(wf["wf_o3_vmr"]).max(axis=0, keepdims=True)
    assert err.max() < 1e-7
