"""Delta-M scaling: Atmosphere::apply_delta_m_scaling, cpp/lib/atmosphere/atmosphere.cpp:69-203,
and its use by the DO layer optics (sktran_do_layerarray.cpp:396-410, 773-800).

Upstream has no stored numbers for it; like the reference (tests/weightingfunctions, src/sasktran2/test_util/wf.py)
the oracle's scaled weighting functions are pinned by central finite differences of the whole chain
(unscaled inputs -> scaling -> solve), the C-ABI host pass is compared with the numpy restatement, and the CUDA path
with the oracle."""
import ctypes as C

import numpy as np
import pytest

from sasktran2_b200 import scenarios as scn
from tests.test_oracle_wf import species_atmosphere


def _scaled_solve(oracle_mod, common, order, k, ssa, leg, albedo, d_leg=None, maps=None, calc_derivs=False):
    sc = oracle_mod.apply_delta_m_scaling(order, ssa, k, leg, d_leg=d_leg, mappings=maps)
    out = oracle_mod.do_radiance(**common, ssa=sc["ssa"], ext=sc["ext"], leg=sc["leg"], albedo=albedo, d_leg=sc["d_leg"],
                                 calc_derivs=calc_derivs, f=sc["f"], d_f=sc["d_f"], stable=True)
    return out, sc


@pytest.mark.parametrize("interp,geotype,nstr", [(2, 0, 4), (2, 1, 8)])
def test_oracle_delta_m_wf_matches_finite_differences(oracle_mod, interp, geotype, nstr):
    nlayers, nwavel, nleg = 7, 2, 3 * nstr
    z, _, k_aer0, k_abs0, w_aer, _ = species_atmosphere(nstr, nlayers, nwavel)
    k_ray = scn.rayleigh_extinction(z)[:, None] * np.logspace(-0.5, 0.7, nwavel)[None, :]
    b_aer, b_ray = scn.hg_moments(0.8, nleg), scn.rayleigh_moments(nleg)   # strongly forward peaked: f ~ 0.8^nstr

    def build(k_aer, k_abs):
        return scn._mix([(k_ray, 1.0, b_ray), (k_aer, w_aer, b_aer), (k_abs, 0.0, np.zeros(nleg))], nleg)

    nloc = z.size
    cz, az, albedo = np.array([0.9, 0.5]), np.array([0.3, 2.1]), 0.25
    common = dict(nstr=nstr, alt=z, interp=interp, geotype=geotype, cos_sza=0.55, los_cos_vza=cz, los_rel_az=az)
    k, ssa, leg, ks = build(k_aer0, k_abs0)
    d_leg = (b_aer[:, None, None] - leg)[..., None]
    maps = {
        "abs": dict(d_extinction=np.ones_like(k), d_ssa=-ssa / k, scat_factor=None, scat_index=-1),
        "aer": dict(d_extinction=np.ones_like(k), d_ssa=(w_aer - ssa) / k, scat_factor=w_aer / ks, scat_index=0),
    }
    base, sc = _scaled_solve(oracle_mod, common, nstr, k, ssa, leg, albedo, d_leg=d_leg, maps=maps, calc_derivs=True)
    assert sc["f"] is not None and sc["f"].max() > 0.05          # the scaling does something here
    unscaled = oracle_mod.do_radiance(**common, ssa=ssa, ext=k, leg=leg, albedo=albedo, stable=True)["radiance"]
    assert np.abs(base["radiance"] / unscaled - 1).max() > 1e-4
    wf = oracle_mod.apply_mappings(base["native"], sc["mappings"], nloc, 1)

    def rad(k_aer, k_abs):
        kk, ss, ll, _ = build(k_aer, k_abs)
        return _scaled_solve(oracle_mod, common, nstr, kk, ss, ll, albedo)[0]["radiance"]

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
            assert np.abs(wf[name][q] - fd).max() / scale.max() < 2e-6, (name, q, wf[name][q], fd)


def test_delta_m_order_not_below_stored_moments_is_a_no_op(oracle_mod):
    k, ssa = np.full((3, 2), 1e-5), np.full((3, 2), 0.9)
    leg = np.ones((4, 3, 2))
    sc = oracle_mod.apply_delta_m_scaling(4, ssa, k, leg)
    assert sc["f"] is None and np.array_equal(sc["leg"], leg) and np.array_equal(sc["ext"], k)


def test_c_abi_delta_m_host_pass_matches_restatement(oracle_mod):
    """sk_atmosphere_apply_delta_m_scaling mutates the caller's arrays and the mappings in place, like upstream."""
    import sasktran2_b200 as sk
    from sasktran2_b200 import _lib

    nstr = 4
    sc = scn.small_wf_case(nstr=nstr, nlayers=9, nwavel=70, nlos=2, nleg=10)   # 70 wavelengths: two host threads
    cfg = sk.Config()
    cfg.num_streams = nstr
    geo = sk.Geometry1D(sc.cos_sza, 0.0, sc.earth_radius, sc.altitudes, sk.InterpolationMethod(sc.interp),
                        sk.GeometryType(sc.geotype))
    atm = sk.Atmosphere.from_scenario(sc, geo, cfg)
    names = sorted(n for n, mp in sc.mappings.items() if "d_legendre" in mp)
    d_leg = np.stack([sc.mappings[n]["d_legendre"] for n in names], axis=-1)
    maps = {n: dict(d_ssa=mp["d_ssa"], d_extinction=mp["d_extinction"], scat_factor=mp.get("scat_factor"),
                    scat_index=names.index(n) if n in names else -1) for n, mp in sc.mappings.items()}
    want = oracle_mod.apply_delta_m_scaling(nstr, sc.ssa, sc.total_extinction, sc.leg_coeff, d_leg=d_leg, mappings=maps)
    cfg.delta_m_scaling = True
    atm.internal_object()          # applies the scaling once
    atm.internal_object()
    assert atm._applied_delta_m_order == nstr
    np.testing.assert_allclose(atm.storage.ssa, want["ssa"], rtol=1e-14)
    np.testing.assert_allclose(atm.storage.total_extinction, want["ext"], rtol=1e-14)
    np.testing.assert_allclose(atm.storage.leg_coeff, want["leg"], rtol=1e-14)
    for n in sc.mappings:
        m = atm.storage.get_derivative_mapping(n)
        np.testing.assert_allclose(m.d_extinction, want["mappings"][n]["d_extinction"], rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(m.d_ssa, want["mappings"][n]["d_ssa"], rtol=1e-12, atol=1e-300)
        if n in names:
            np.testing.assert_allclose(m.d_leg_coeff, want["d_leg"][..., names.index(n)], rtol=1e-12, atol=1e-14)
    # like upstream, the C entry point rescales whatever the storage holds (the caller refills between applications;
    # tests/test_host_logic.py covers zero_storage + refill): after set_zero the storage is unscaled again
    assert _lib.lib().sk_atmosphere_storage_set_zero(atm.storage._h) == 0
    assert np.all(atm.storage.leg_coeff == 0)


@pytest.mark.gpu
@pytest.mark.parametrize("nstr,nleg,geotype", [(4, 12, 0), (8, 24, 1), (16, 32, 1)])
def test_cuda_delta_m_vs_oracle(oracle_mod, nstr, nleg, geotype):
    """Radiances (1e-9) and weighting functions (1e-7 of the column maximum; the 1/k-amplified scatterer mapping at
    its noise floor, see tests/test_gpu_parity.py::_assert_wf) with delta-M scaling applied through the C ABI."""
    import sasktran2_b200 as sk
    from tests.test_gpu_parity import _assert_wf

    sc = scn.small_wf_case(nstr=nstr, nlayers=10, nwavel=4, nlos=3, geotype=geotype, nleg=nleg)
    cfg, geo, view, eng, atm = sk.engine_for_scenario(sc)
    cfg.delta_m_scaling = True
    atm.surface.enable_albedo_derivative("wf_albedo")
    res = eng.calculate_radiance(atm)
    # the oracle solves the scaled problem: scenario arrays replaced by the restatement's scaled copies
    names = sorted(n for n, mp in sc.mappings.items() if "d_legendre" in mp)
    d_leg = np.stack([sc.mappings[n]["d_legendre"] for n in names], axis=-1)
    maps = {n: dict(d_ssa=mp["d_ssa"], d_extinction=mp["d_extinction"], scat_factor=mp.get("scat_factor"),
                    scat_index=names.index(n) if n in names else -1) for n, mp in sc.mappings.items()}
    want = oracle_mod.apply_delta_m_scaling(nstr, sc.ssa, sc.total_extinction, sc.leg_coeff, d_leg=d_leg, mappings=maps)
    unscaled = oracle_mod.do_radiance(nstr=sc.nstr, alt=sc.altitudes, interp=sc.interp, geotype=sc.geotype,
                                      cos_sza=sc.cos_sza, earth_radius=sc.earth_radius, los_cos_vza=sc.los_cos_vza,
                                      los_rel_az=sc.los_rel_az, ssa=sc.ssa, ext=sc.total_extinction, leg=sc.leg_coeff,
                                      albedo=sc.albedo)["radiance"]
    assert np.abs(res["radiance"][:, :, 0] / unscaled - 1).max() > 1e-6   # the scaling changed the answer
    sc.ssa, sc.total_extinction, sc.leg_coeff = want["ssa"], want["ext"], want["leg"]
    for n, mp in sc.mappings.items():
        mp["d_ssa"], mp["d_extinction"] = want["mappings"][n]["d_ssa"], want["mappings"][n]["d_extinction"]
        if n in names:
            mp["d_legendre"] = want["d_leg"][..., names.index(n)]
    sc.delta_m = dict(f=want["f"], d_f=want["d_f"])
    _assert_wf(oracle_mod, sc, res)


@pytest.mark.gpu
def test_cuda_twostream_source_reference_case(oracle_mod):
    """The inputs of the reference's test_twostream_delta_m_matches_two_stream_discrete_ordinates
    (tests/engine/test_twostream.py:104-160: plane parallel, 8 layers, 9 wavelengths, 4 stored moments, delta-M,
    single scatter off, two ground-viewing rays), run with multiple_scatter_source = TwoStream and with
    DiscreteOrdinates (2 streams): both must equal the oracle's multiple-scatter-only 2-stream solve, radiances to
    1e-9 and the extinction / single-scatter-albedo weighting functions to 1e-7 of the column maximum (upstream asserts
    2e-8 between its two implementations)."""
    import sasktran2_b200 as sk

    z = np.arange(0.0, 40_001.0, 5_000.0)
    nw, nloc = 9, z.size
    spectral = 0.8 + 0.04 * np.arange(nw)[None, :]
    k = (2.0e-5 * np.exp(-z[:, None] / 8_000.0) + 1.0e-8) * spectral
    ssa = np.full((nloc, nw), 0.87)
    g = 0.62 + 0.01 * np.arange(nw)[None, :] / nw
    leg = np.zeros((4, nloc, nw))
    leg[0], leg[1], leg[2], leg[3] = 1.0, 3.0 * g, 5.0 * g**2, 7.0 * g**3
    cz, az = np.array([0.7, 0.35]), np.array([0.3, -0.4])
    ones, zeros = np.ones((nloc, nw)), np.zeros((nloc, nw))
    maps = {"wf_extinction": dict(d_extinction=ones, d_ssa=zeros, scat_factor=None, scat_index=-1),
            "wf_ssa": dict(d_extinction=zeros, d_ssa=ones, scat_factor=None, scat_index=-1)}
    want = oracle_mod.apply_delta_m_scaling(2, ssa, k, leg, mappings=maps)
    ora = oracle_mod.do_radiance(nstr=2, alt=z, interp=1, geotype=0, cos_sza=0.6, earth_radius=6_371_000.0,
                                 los_cos_vza=cz, los_rel_az=az, ssa=want["ssa"], ext=want["ext"], leg=want["leg"],
                                 solar=np.full(nw, 1.1), albedo=0.2, include_ss=False, calc_derivs=True, f=want["f"],
                                 stable=True)
    wf = oracle_mod.apply_mappings(ora["native"], want["mappings"], nloc, 0)
    results = {}
    for source in (sk.MultipleScatterSource.TwoStream, sk.MultipleScatterSource.DiscreteOrdinates):
        cfg = sk.Config()
        cfg.num_streams = 2
        cfg.num_stokes = 1
        cfg.do_backprop = True
        cfg.delta_m_scaling = True
        cfg.single_scatter_source = sk.SingleScatterSource.NoSource
        cfg.multiple_scatter_source = source
        geo = sk.Geometry1D(0.6, 0.2, 6_371_000.0, z, sk.InterpolationMethod.LinearInterpolation,
                            sk.GeometryType.PlaneParallel)
        view = sk.ViewingGeometry()
        for c, a in zip(cz, az):
            view.add_ray(sk.GroundViewingSolar(0.6, float(a), float(c), 200_000.0))
        atm = sk.Atmosphere(geo, cfg, numwavel=nw, calculate_derivatives=True, num_legendre=4)
        atm.storage.total_extinction[:] = k
        atm.storage.ssa[:] = ssa
        atm.storage.leg_coeff[:] = leg
        atm.storage.solar_irradiance[:] = 1.1
        atm.surface.albedo[:] = 0.2
        for n, mp in maps.items():
            m = atm.storage.get_derivative_mapping(n)
            m.d_extinction[:] = mp["d_extinction"]
            m.d_ssa[:] = mp["d_ssa"]
        res = sk.Engine(cfg, geo, view).calculate_radiance(atm)
        np.testing.assert_allclose(res["radiance"][:, :, 0], ora["radiance"], rtol=1e-9)
        for n in maps:
            err = np.abs(res[n][..., 0] - wf[n]) / np.abs(wf[n]).max(axis=0, keepdims=True)
            assert err.max() < 1e-7, (source, n, float(err.max()))
        results[source] = res
    a, b = results[sk.MultipleScatterSource.TwoStream], results[sk.MultipleScatterSource.DiscreteOrdinates]
    for n in a:
        if not n.startswith("_"):
            np.testing.assert_array_equal(a[n], b[n])


@pytest.mark.gpu
@pytest.mark.parametrize("geotype,nlos", [(0, 2), (1, 2), (1, 5), (0, 1)])
def test_cuda_dedicated_twostream_kernel_vs_oracle(oracle_mod, geotype, nlos):
    """multiple_scatter_source = TwoStream without weighting functions runs the dedicated single-sweep kernel
    (k_twostream, one launch per chunk): radiances against the restatement of the reference's two-stream source
    (oracle/twostream_oracle.hpp, 1e-9) on a line-by-line spectrum of the configs[2] shape (60 layers, optical depths
    over nine decades) and on the reference test's delta-M case."""
    import sasktran2_b200 as sk
    from tests.test_oracle_twostream import reference_case

    def run(alt, interp, geotype, cos_sza, los_cos_vza, los_rel_az, ssa, ext, leg, albedo, solar=None, earth_radius=6372000.0,
            delta_m=False):
        cfg = sk.Config()
        cfg.num_streams = 2
        cfg.delta_m_scaling = delta_m
        cfg.single_scatter_source = sk.SingleScatterSource.NoSource
        cfg.multiple_scatter_source = sk.MultipleScatterSource.TwoStream
        geo = sk.Geometry1D(cos_sza, 0.0, earth_radius, alt, sk.InterpolationMethod(interp), sk.GeometryType(geotype))
        view = sk.ViewingGeometry()
        for c, a in zip(los_cos_vza, los_rel_az):
            view.add_ray(sk.GroundViewingSolar(cos_sza, float(a), float(c), 200_000.0))
        nw = ssa.shape[1]
        atm = sk.Atmosphere(geo, cfg, numwavel=nw, calculate_derivatives=False, num_legendre=leg.shape[0])
        atm.storage.total_extinction[:] = ext
        atm.storage.ssa[:] = ssa
        atm.storage.leg_coeff[:] = leg
        atm.surface.albedo[:] = albedo
        if solar is not None:
            atm.storage.solar_irradiance[:] = solar
        eng = sk.Engine(cfg, geo, view)
        rad = eng.calculate_radiance(atm)["radiance"][:, :, 0].copy()
        assert eng.kernel_launches() == 2          # input validation + the dedicated kernel, not the discrete-ordinates pipeline
        return rad

    c3 = scn.config3(nwavel=20000, nlayers=60, nlos=nlos)
    c3.geotype = geotype
    inp = dict(alt=c3.altitudes, interp=c3.interp, geotype=c3.geotype, cos_sza=c3.cos_sza, los_cos_vza=c3.los_cos_vza,
               los_rel_az=c3.los_rel_az, ssa=c3.ssa, ext=c3.total_extinction, leg=c3.leg_coeff, albedo=c3.albedo)
    np.testing.assert_allclose(run(**inp), oracle_mod.twostream_radiance(**inp)["radiance"], rtol=1e-9)
    geo, ssa, k, leg = reference_case()
    geo["geotype"] = geotype
    sc = oracle_mod.apply_delta_m_scaling(2, ssa, k, leg)
    want = oracle_mod.twostream_radiance(**geo, ssa=sc["ssa"], ext=sc["ext"], leg=sc["leg"], f=sc["f"])["radiance"]
    np.testing.assert_allclose(run(**geo, ssa=ssa, ext=k, leg=leg, delta_m=True), want, rtol=1e-9)
