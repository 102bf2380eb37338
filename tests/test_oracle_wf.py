"""Weighting functions of the oracle vs central finite differences, as the reference validates its own
(src/sasktran2/test_util/wf.py:9-80; tests/weightingfunctions/test_lowlevel.py)."""
import numpy as np
import pytest

from sasktran2_b200 import scenarios as scn


def species_atmosphere(nstr, nlayers, nwavel, scale=None):
    """Rayleigh + aerosol + absorber; returns builder(k_aer, k_abs) -> (k, ssa, leg) and mapping pieces."""
    nleg = nstr
    z = np.linspace(0.0, 60e3, nlayers + 1)
    s = np.logspace(-0.5, 0.7, nwavel)
    k_ray = scn.rayleigh_extinction(z)[:, None] * s[None, :]
    k_aer0 = (2e-5 * np.exp(-z / 4e3))[:, None] * np.ones(nwavel)[None, :]
    k_abs0 = (1.5e-5 * np.exp(-(((z - 22e3) / 9e3) ** 2)))[:, None] * np.linspace(0.3, 1.0, nwavel)[None, :]
    w_aer = 0.93
    b_aer = scn.hg_moments(0.65, nleg)
    b_ray = scn.rayleigh_moments(nleg)

    def build(k_aer, k_abs):
        k, ssa, leg, ks = scn._mix([(k_ray, 1.0, b_ray), (k_aer, w_aer, b_aer), (k_abs, 0.0, np.zeros(nleg))], nleg)
        return k, ssa, leg, ks

    return z, build, k_aer0, k_abs0, w_aer, b_aer


@pytest.mark.parametrize("stable", [False, True], ids=["reference-formulas", "stable-multipliers"])
@pytest.mark.parametrize("interp,geotype,nstr", [(2, 0, 4), (2, 1, 8), (1, 1, 8)])
def test_oracle_wf_matches_finite_differences(oracle_mod, interp, geotype, nstr, stable):
    nlayers, nwavel = 8, 2
    z, build, k_aer0, k_abs0, w_aer, b_aer = species_atmosphere(nstr, nlayers, nwavel)
    nloc = z.size
    cz = np.array([0.9, 0.5])
    az = np.array([0.3, 2.1])
    albedo = 0.25
    common = dict(nstr=nstr, alt=z, interp=interp, geotype=geotype, cos_sza=0.55, los_cos_vza=cz, los_rel_az=az)

    k, ssa, leg, ks = build(k_aer0, k_abs0)
    d_leg = (b_aer[:, None, None] - leg)[..., None]
    base = oracle_mod.do_radiance(**common, ssa=ssa, ext=k, leg=leg, albedo=albedo, d_leg=d_leg, calc_derivs=True,
                                  stable=stable)
    maps = {
        "abs": dict(d_extinction=np.ones_like(k), d_ssa=-ssa / k),
        "aer": dict(d_extinction=np.ones_like(k), d_ssa=(w_aer - ssa) / k, scat_factor=w_aer / ks, scat_index=0),
    }
    wf = oracle_mod.apply_mappings(base["native"], maps, nloc, 1)
    d_albedo = base["native"][:, :, -1]

    def rad(k_aer, k_abs, alb=albedo):
        kk, ss, ll, _ = build(k_aer, k_abs)
        return oracle_mod.do_radiance(**common, ssa=ss, ext=kk, leg=ll, albedo=alb, stable=stable)["radiance"]

    # with linear interpolation the reference takes the scattering-derivative direction from one grid point
    # only (sktran_do_layerarray.cpp:761-800, see oracle map_to_native), so its aerosol WF is approximate there.
    # `lower` interpolation is exact; on the grid's top point `lower` never contributes (weight 0).
    for q in (0, 3, nlayers - 1):
        for name, which in (("abs", 1), ("aer", 0)):
            fd = np.zeros((nwavel, cz.size))
            for w in range(nwavel):
                pert = [k_aer0.copy(), k_abs0.copy()]
                h = 1e-4 * k[q, w]  # relative to the total extinction: keeps FD rounding noise small
                pert[which][q, w] += h
                up = rad(*pert)[w]
                pert[which][q, w] -= 2 * h
                dn = rad(*pert)[w]
                fd[w] = (up - dn) / (2 * h)
            an = wf[name][q]
            scale = np.abs(wf[name]).max(axis=0)
            tol = 2e-6 if (interp == 2 or name == "abs") else 5e-3
            assert np.abs(an - fd).max() / scale.max() < tol, (name, q, an, fd)
    h = 1e-5
    fd = (rad(k_aer0, k_abs0, albedo + h) - rad(k_aer0, k_abs0, albedo - h)) / (2 * h)
    np.testing.assert_allclose(d_albedo, fd, rtol=2e-7)


def test_oracle_stable_multipliers_agree_with_reference_formulas(oracle_mod):
    """The singularity-free multipliers (phi/psi forms) reproduce the reference's formulas to rounding on the
    radiance and to within the reference formulas' own rounding noise on the native derivatives; their own noise
    floor on dI/dk is orders of magnitude lower (16 streams at cos_sza = 0.6: some layer always sits close to the
    secant = eigenvalue degeneracy, DESIGN.md "Conditioning")."""
    sc = scn.small_wf_case(nstr=16, nlayers=25, nwavel=5, nlos=6)
    d_leg = sc.mappings["wf_aerosol_extinction"]["d_legendre"][..., None]
    inp = dict(nstr=sc.nstr, alt=sc.altitudes, interp=sc.interp, geotype=sc.geotype, cos_sza=sc.cos_sza,
               los_cos_vza=sc.los_cos_vza, los_rel_az=sc.los_rel_az, ssa=sc.ssa, leg=sc.leg_coeff, albedo=sc.albedo,
               d_leg=d_leg, calc_derivs=True)
    nloc = sc.nloc
    k_block = slice(0, nloc)

    def run(stable, eps=0.0):
        return oracle_mod.do_radiance(**inp, ext=sc.total_extinction * (1.0 + eps), stable=stable)

    ref, stb = run(False), run(True)
    np.testing.assert_allclose(stb["radiance"], ref["radiance"], rtol=1e-12)

    def rel(a, b):
        return np.max(np.abs(a - b)[..., k_block] / np.abs(b[..., k_block]).max(axis=-1, keepdims=True))

    noise_ref = max(rel(run(False, e)["native"], ref["native"]) for e in (1e-12, -1e-12, 1e-11))
    noise_stb = max(rel(run(True, e)["native"], stb["native"]) for e in (1e-12, -1e-12, 1e-11))
    assert rel(stb["native"], ref["native"]) <= 10.0 * max(noise_ref, noise_stb)
    assert noise_stb < 1e-9


@pytest.mark.parametrize("nstr,nlayers,nlos,interp,geotype", [(2, 9, 2, 1, 1), (4, 9, 2, 2, 0), (8, 12, 3, 1, 1), (16, 25, 4, 1, 1),
                                                              (8, 1, 2, 1, 1)])
def test_oracle_reverse_mode_matches_forward_mode(oracle_mod, nstr, nlayers, nlos, interp, geotype):
    """The reverse-mode linearisation (RTESolver::backprop, sktran_do_rte.cpp:1793-1895: layer-local duals, one
    transposed band solve per line of sight, cross-layer chain of beam transmittance and secant) gives the layer lanes
    and native derivatives of the dense forward-mode duals, with one and two scattering groups and the albedo lane.
    Singularity-free multipliers: the two differentiation orders of the reference's direct C+ / D- formulas differ by
    their own rounding noise (1e-6 of the column maximum at 16 streams, DESIGN.md "Conditioning")."""
    sc = scn.small_wf_case(nstr=nstr, nlayers=nlayers, nwavel=3, nlos=nlos, interp=interp, geotype=geotype)
    aer = sc.mappings["wf_aerosol_extinction"]["d_legendre"]
    for d_leg in (None, aer[..., None], np.stack([aer, 0.5 * aer + 0.1], axis=-1)):
        kw = dict(nstr=sc.nstr, alt=sc.altitudes, interp=sc.interp, geotype=sc.geotype, cos_sza=sc.cos_sza,
                  earth_radius=sc.earth_radius, los_cos_vza=sc.los_cos_vza, los_rel_az=sc.los_rel_az, ssa=sc.ssa,
                  ext=sc.total_extinction, leg=sc.leg_coeff, albedo=sc.albedo, d_leg=d_leg, calc_derivs=True,
                  return_lanes=True, stable=True)
        fwd = oracle_mod.do_radiance(**kw)
        rev = oracle_mod.do_radiance(**kw, reverse=True)
        np.testing.assert_allclose(rev["radiance"], fwd["radiance"], rtol=1e-12)
        for key in ("lanes", "native"):
            scale = np.abs(fwd[key]).max(axis=2, keepdims=True)
            assert np.max(np.abs(rev[key] - fwd[key]) / scale) < 1e-9, key
