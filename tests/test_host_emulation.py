"""Runs the product's kernel bodies on the CPU (tests/host_emul.cpp) and compares with the oracle and the
reference's DISORT tables.  This is a test harness, not a fallback: the shipped library has no CPU path."""
import ctypes
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

from tests.test_oracle_golden import GOLD, case_inputs

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def emul():
    so = ROOT / "tests" / "libhost_emul.so"
    srcs = [ROOT / "tests" / "host_emul.cpp", ROOT / "sasktran2_b200" / "csrc" / "disco_plan.cpp",
            ROOT / "sasktran2_b200" / "csrc" / "disco_brdf.cpp",
            ROOT / "sasktran2_b200" / "csrc" / "disco_core.h", ROOT / "sasktran2_b200" / "csrc" / "disco_bodies.h",
            ROOT / "sasktran2_b200" / "csrc" / "disco_bvp_rows.h", ROOT / "sasktran2_b200" / "csrc" / "disco_wf_body.h",
            ROOT / "sasktran2_b200" / "csrc" / "disco_twostream_body.h"]
    if not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", str(so), str(srcs[0]), str(srcs[1]), str(srcs[2])],
                       check=True)
    lib = ctypes.CDLL(str(so))
    lib.emul_last_error.restype = ctypes.c_char_p

    def P(a):
        return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))

    def run(nstr, alt, interp, geotype, cos_sza, los_cos_vza, los_rel_az, ssa, ext, leg, albedo,
            earth_radius=6372000.0, include_ss=True, solar=None, d_leg=None, want_native=False, emission=None,
            surface_emission=None, modis_args=None, **_):
        alt = np.ascontiguousarray(alt, float)
        ssa = np.asfortranarray(ssa, float)
        ext = np.asfortranarray(ext, float)
        leg = np.asfortranarray(leg, float)
        nloc, nw, nleg = alt.size, ssa.shape[1], leg.shape[0]
        cz = np.ascontiguousarray(los_cos_vza, float)
        az = np.ascontiguousarray(los_rel_az, float)
        solar = np.ones(nw) if solar is None else np.ascontiguousarray(solar, float)
        alb = np.ascontiguousarray(np.broadcast_to(albedo, (nw,)), float)
        rad = np.zeros((nw, cz.size))
        naz = ctypes.c_int(0)
        G = 0
        dl = None
        if d_leg is not None:
            dl = np.asfortranarray(d_leg, float)  # [nleg, nloc, nw, G]
            G = dl.shape[3]
        native = np.zeros((nw, cz.size, nloc * (2 + G) + 1)) if want_native else None
        em = None if emission is None else np.asfortranarray(emission, float)
        se = None if surface_emission is None else np.ascontiguousarray(np.broadcast_to(surface_emission, (nw,)), float)
        lib.emul_set_emission(P(em) if em is not None else None, P(se) if se is not None else None)
        ma = None if modis_args is None else np.asfortranarray(modis_args, float)   # [3, nwavel]
        lib.emul_set_modis(P(ma) if ma is not None else None)
        rc = lib.emul_do_radiance(nstr, nloc, nw, nleg, cz.size, P(alt), interp, geotype, ctypes.c_double(cos_sza),
                                  ctypes.c_double(earth_radius), P(cz), P(az), P(ssa), P(ext), P(leg), P(solar), P(alb),
                                  int(include_ss), P(rad), ctypes.byref(naz), P(dl) if dl is not None else None, G,
                                  P(native) if native is not None else None)
        lib.emul_set_emission(None, None)
        lib.emul_set_modis(None)
        if rc:
            raise RuntimeError(lib.emul_last_error().decode())
        if want_native:
            return rad, naz.value, native
        return rad, naz.value

    def run_twostream(alt, interp, geotype, cos_sza, los_cos_vza, los_rel_az, ssa, ext, leg, albedo,
                      earth_radius=6372000.0, solar=None, f=None, **_):
        alt = np.ascontiguousarray(alt, float)
        ssa = np.asfortranarray(ssa, float)
        ext = np.asfortranarray(ext, float)
        leg = np.asfortranarray(leg, float)
        nloc, nw, nleg = alt.size, ssa.shape[1], leg.shape[0]
        cz = np.ascontiguousarray(los_cos_vza, float)
        az = np.ascontiguousarray(los_rel_az, float)
        solar = np.ones(nw) if solar is None else np.ascontiguousarray(solar, float)
        alb = np.ascontiguousarray(np.broadcast_to(albedo, (nw,)), float)
        fd = None if f is None else np.asfortranarray(f, float)
        rad = np.zeros((nw, cz.size))
        rc = lib.emul_twostream(nloc, nw, nleg, cz.size, P(alt), interp, geotype, ctypes.c_double(cos_sza),
                                ctypes.c_double(earth_radius), P(cz), P(az), P(ssa), P(ext), P(leg), P(solar), P(alb),
                                P(fd) if fd is not None else None, P(rad))
        if rc:
            raise RuntimeError(lib.emul_last_error().decode())
        return rad

    run.twostream = run_twostream
    return run


@pytest.mark.parametrize("case", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_kernel_bodies_match_disort_tables_and_oracle(emul, oracle_mod, case):
    inp = case_inputs(case)
    rad, _ = emul(**inp)
    np.testing.assert_allclose(rad[0] * case["sun"]["direct"], np.array(case["radiance"]), rtol=0, atol=case["abs_tol"])
    ora = oracle_mod.do_radiance(**inp)["radiance"]
    # conservative layers (SSA dithered to 1 - 1e-9) carry O(1e-7) cancellation noise in S+ X in BOTH
    # implementations (the reference relaxes its own tolerance to 1e-6 there, test_scalar.cpp:152-166)
    rtol = 1e-6 if case["name"] == "SSA = 1" else 1e-11
    np.testing.assert_allclose(rad, ora, rtol=rtol, atol=0)


def test_kernel_bodies_pseudo_spherical_and_nadir_skip(emul, oracle_mod):
    from sasktran2_b200 import scenarios

    sc = scenarios.config2(nwavel=3, nlayers=20, nstr=8, nlos=4)
    inp = dict(nstr=sc.nstr, alt=sc.altitudes, interp=sc.interp, geotype=sc.geotype, cos_sza=sc.cos_sza,
               los_cos_vza=sc.los_cos_vza, los_rel_az=sc.los_rel_az, ssa=sc.ssa, ext=sc.total_extinction,
               leg=sc.leg_coeff, albedo=sc.albedo)
    rad, naz = emul(**inp)
    assert naz == 8
    np.testing.assert_allclose(rad, oracle_mod.do_radiance(**inp)["radiance"], rtol=1e-10)
    sc1 = scenarios.config1(nwavel=5, nlayers=10)
    inp = dict(nstr=sc1.nstr, alt=sc1.altitudes, interp=sc1.interp, geotype=sc1.geotype, cos_sza=sc1.cos_sza,
               los_cos_vza=sc1.los_cos_vza, los_rel_az=sc1.los_rel_az, ssa=sc1.ssa, ext=sc1.total_extinction,
               leg=sc1.leg_coeff, albedo=sc1.albedo)
    rad, naz = emul(**inp)
    assert naz == 1  # exactly-nadir LOS: azimuth orders m > 0 contribute exactly zero and are skipped
    np.testing.assert_allclose(rad, oracle_mod.do_radiance(**inp)["radiance"], rtol=1e-10)


@pytest.mark.parametrize("nstr,interp,geotype,nlos", [(4, 2, 0, 2), (8, 1, 1, 3), (8, 1, 1, 6), (16, 1, 1, 2), (2, 1, 1, 2)])
def test_kernel_bodies_weighting_functions_match_oracle(emul, oracle_mod, nstr, interp, geotype, nlos):
    """Reverse-mode (adjoint BVP + layer-local duals + cross-layer chain) native derivatives of the product's
    kernel bodies vs the oracle's forward-mode duals.  Tolerance: 1e-7 relative to the largest derivative of
    each kind (BASELINE.json north_star: 1e-7 on weighting functions)."""
    from sasktran2_b200 import scenarios

    sc = scenarios.small_wf_case(nstr=nstr, nlayers=9, nwavel=2, nlos=nlos, interp=interp, geotype=geotype)
    d_leg = sc.mappings["wf_aerosol_extinction"]["d_legendre"][..., None]
    inp = dict(nstr=sc.nstr, alt=sc.altitudes, interp=sc.interp, geotype=sc.geotype, cos_sza=sc.cos_sza,
               los_cos_vza=sc.los_cos_vza, los_rel_az=sc.los_rel_az, ssa=sc.ssa, ext=sc.total_extinction,
               leg=sc.leg_coeff, albedo=sc.albedo)
    rad, _, native = emul(**inp, d_leg=d_leg, want_native=True)
    ora = oracle_mod.do_radiance(**inp, d_leg=d_leg, calc_derivs=True)
    np.testing.assert_allclose(rad, ora["radiance"], rtol=1e-10)
    nloc = sc.nloc
    for lo, hi in ((0, nloc), (nloc, 2 * nloc), (2 * nloc, 3 * nloc), (3 * nloc, 3 * nloc + 1)):
        a, b = native[..., lo:hi], ora["native"][..., lo:hi]
        scale = np.abs(b).max(axis=-1, keepdims=True)
        assert np.max(np.abs(a - b) / scale) < 1e-7, (lo, np.max(np.abs(a - b) / scale))


@pytest.mark.parametrize("nstr,nlayers,nlos", [(8, 20, 3), (16, 40, 4)])
def test_kernel_bodies_pass_the_weighting_function_parity_rules(emul, oracle_mod, nstr, nlayers, nlos):
    """The parity rules the GPU tests apply to the CUDA results (tests/wf_checks.py: flat 1e-7 against the stable and the
    reference-formula oracle, explicit list of optically thin grid points for the extinction probe) applied to the
    product's kernel bodies run on the host, on an atmosphere that reaches 100 km (layer optical depths down to 1e-8)."""
    from sasktran2_b200 import scenarios
    from tests import wf_checks

    sc = scenarios.small_wf_case(nstr=nstr, nlayers=nlayers, nwavel=4, nlos=nlos)
    wf_checks.add_native_probes(sc, scat_probe=True)
    names = wf_checks.scat_names(sc)
    d_leg = np.stack([sc.mappings[n]["d_legendre"] for n in names], axis=-1)
    inp = dict(nstr=sc.nstr, alt=sc.altitudes, interp=sc.interp, geotype=sc.geotype, cos_sza=sc.cos_sza,
               los_cos_vza=sc.los_cos_vza, los_rel_az=sc.los_rel_az, ssa=sc.ssa, ext=sc.total_extinction,
               leg=sc.leg_coeff, albedo=sc.albedo)
    rad, _, native = emul(**inp, d_leg=d_leg, want_native=True)
    res = wf_checks.candidate_from_native(oracle_mod, sc, rad, native)
    rep = wf_checks.assert_wf(oracle_mod, sc, res)
    assert rep["reference"]["wf_o3_vmr"]["rule"] == "flat 1e-7"
    assert rep["reference"]["wf_probe_k"]["listed_points"]["count"] > 0


@pytest.mark.parametrize("geotype,nlos", [(0, 2), (1, 2), (1, 3), (1, 1)])
def test_twostream_kernel_body_matches_the_two_stream_oracle(emul, oracle_mod, geotype, nlos):
    """The single-sweep two-stream kernel body ((U^-T w).z instead of a back substitution, disco_twostream_body.h) against
    the restatement of the reference's two-stream source (oracle/twostream_oracle.hpp): the reference test's inputs
    (delta-M scaled), an O2-A-band-like line-by-line case with optical depths over nine decades, thin resonant layers."""
    from sasktran2_b200 import scenarios
    from tests.test_oracle_twostream import reference_case

    geo, ssa, k, leg = reference_case()
    geo["geotype"] = geotype
    geo["los_cos_vza"] = np.array([0.7, 0.35, 1.0])[:nlos]
    geo["los_rel_az"] = np.array([0.3, -0.4, 0.0])[:nlos]
    sc = oracle_mod.apply_delta_m_scaling(2, ssa, k, leg)
    got = emul.twostream(**geo, ssa=sc["ssa"], ext=sc["ext"], leg=sc["leg"], f=sc["f"])
    want = oracle_mod.twostream_radiance(**geo, ssa=sc["ssa"], ext=sc["ext"], leg=sc["leg"], f=sc["f"])["radiance"]
    np.testing.assert_allclose(got, want, rtol=1e-11)
    thin = k.copy()
    thin[-3:] *= 1e-9
    got = emul.twostream(**geo, ssa=ssa, ext=thin, leg=leg)
    want = oracle_mod.twostream_radiance(**geo, ssa=ssa, ext=thin, leg=leg)["radiance"]
    np.testing.assert_allclose(got, want, rtol=1e-10)
    c3 = scenarios.config3(nwavel=200, nlayers=60, nlos=nlos)
    c3.geotype = geotype
    inp = dict(alt=c3.altitudes, interp=c3.interp, geotype=c3.geotype, cos_sza=c3.cos_sza, los_cos_vza=c3.los_cos_vza,
               los_rel_az=c3.los_rel_az, ssa=c3.ssa, ext=c3.total_extinction, leg=c3.leg_coeff, albedo=c3.albedo)
    got = emul.twostream(**inp)
    want = oracle_mod.twostream_radiance(**inp)["radiance"]
    assert np.all(want > 0)
    np.testing.assert_allclose(got, want, rtol=1e-9)
