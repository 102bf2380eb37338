"""Pins the CPU oracle to the reference's own DISORT-verified known-answer tables
(cpp/lib/tests/sktran_disco/legacy/test_scalar.cpp:84-878, tolerance SKDO_FPC_EPS = 1e-8, `:19`)."""
import json
from pathlib import Path

import numpy as np
import pytest

GOLD = json.loads((Path(__file__).parent / "golden" / "disort_scalar.json").read_text())


def case_inputs(c):
    """Build the reference test driver's atmosphere (cpp/lib/tests/sktran_disco/test_util.cpp:26-115)."""
    nstr = c["nstr"]
    layers = c["layers_od_ssa_g"]
    nl = len(layers)
    nloc = nl + 1
    alt = np.arange(nloc, dtype=float)
    ext = np.zeros((nloc, 1))
    ssa = np.zeros((nloc, 1))
    leg = np.zeros((nstr, nloc, 1))
    for l, (od, w, g) in enumerate(layers):
        q = nl - l - 1
        ext[q, 0] = od
        ssa[q, 0] = w
        leg[:, q, 0] = [(2 * k + 1) * g**k for k in range(nstr)]
    cz = np.array([x[0] for x in c["los_coszen_az"]])
    az = np.array([(x[1][1] * np.pi / x[1][2]) if x[1][0] == "pi_frac" else x[1][1] for x in c["los_coszen_az"]])
    az = -c["sun"]["saz"] + az
    return dict(nstr=nstr, alt=alt, interp=2, geotype=0, cos_sza=c["sun"]["csz"], los_cos_vza=cz, los_rel_az=az,
                ssa=ssa, ext=ext, leg=leg, albedo=c["albedo"])


@pytest.mark.parametrize("case", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_oracle_matches_disort_tables(oracle_mod, case):
    r = oracle_mod.do_radiance(**case_inputs(case))
    rad = r["radiance"][0] * case["sun"]["direct"]
    np.testing.assert_allclose(rad, np.array(case["radiance"]), rtol=0, atol=case["abs_tol"])


def test_oracle_boundary_conditions(oracle_mod):
    """'Scalar Boundary Conditions' (test_scalar.cpp:847-880): three identical conservative Rayleigh layers
    must equal one layer of 3x the optical depth."""
    nstr = 16
    cz = np.repeat([1.0, 0.8, 0.6, 0.4, 0.2], 7)
    az = np.tile(np.arange(7) * np.pi / 6, 5)

    def run(ods):
        nl = len(ods)
        alt = np.arange(nl + 1, dtype=float)
        ext = np.zeros((nl + 1, 1))
        ssa = np.zeros((nl + 1, 1))
        leg = np.zeros((nstr, nl + 1, 1))
        for l, od in enumerate(ods):
            q = nl - l - 1
            ext[q, 0] = od
            ssa[q, 0] = 1.0
            leg[0, q, 0] = 1.0
            leg[2, q, 0] = 0.5
        return oracle_mod.do_radiance(nstr=nstr, alt=alt, interp=2, geotype=0, cos_sza=0.8, los_cos_vza=cz,
                                      los_rel_az=az, ssa=ssa, ext=ext, leg=leg, albedo=0.8)["radiance"][0]

    np.testing.assert_allclose(run([0.2, 0.2, 0.2]), run([0.6]), rtol=0, atol=1e-8)


def test_band_solver_known_solutions(oracle_mod):
    """Band LU (dgbtf2/dgbtrs restatement) on synthetic band systems with the analytic entries used by the
    reference's generator sin(0.13 (r+1)(c+1)) (cpp/lib/tests/sktran_disco/test_band_factorization.cpp:64-227):
    forward and transposed solves against numpy dense solves, 2e-13 like the reference."""
    rng = np.random.default_rng(0)
    for n, kl in [(12, 2), (40, 5), (96, 11), (160, 23)]:
        r, c = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
        a = np.sin(0.13 * (r + 1) * (c + 1)) + 0.05 * rng.standard_normal((n, n))
        a[np.abs(r - c) > kl] = 0.0
        a += np.diag(0.3 * np.ones(n))  # keep it comfortably non-singular but not diagonally dominant
        xs = np.sin(0.13 * (np.arange(n) + 1))
        for trans in (False, True):
            b = (a.T if trans else a) @ xs
            x = oracle_mod.band_solve(a, b, kl, trans=trans)
            ref = np.linalg.solve(a.T if trans else a, b)
            scale = np.abs(ref).max()
            assert np.abs(x - ref).max() / scale < 2e-11 * max(1.0, np.linalg.cond(a) * 1e-3)
