"""Pins the restatement of the reference's dedicated two-stream source (oracle/twostream_oracle.hpp).

The reference asserts its two-stream source equal to its two-stream discrete-ordinates source with single scatter off,
rtol 2e-8 / atol 2e-12 (tests/engine/test_twostream.py:104-160).  The same inputs go through the restatement and
through the discrete-ordinates oracle (nstr = 2, include_ss = False), which is pinned to the reference's DISORT tables."""
import numpy as np
import pytest


def reference_case(nw=9):
    z = np.arange(0.0, 40_001.0, 5_000.0)
    nloc = z.size
    spectral = 0.8 + 0.04 * np.arange(nw)[None, :]
    k = (2.0e-5 * np.exp(-z[:, None] / 8_000.0) + 1.0e-8) * spectral
    ssa = np.full((nloc, nw), 0.87)
    g = 0.62 + 0.01 * np.arange(nw)[None, :] / nw
    leg = np.zeros((4, nloc, nw))
    leg[0], leg[1], leg[2], leg[3] = 1.0, 3.0 * g, 5.0 * g**2, 7.0 * g**3
    return dict(alt=z, interp=1, geotype=0, cos_sza=0.6, earth_radius=6_371_000.0, los_cos_vza=np.array([0.7, 0.35]),
                los_rel_az=np.array([0.3, -0.4]), solar=np.full(nw, 1.1), albedo=0.2), ssa, k, leg


def test_twostream_source_equals_two_stream_do_on_the_reference_test_inputs(oracle_mod):
    geo, ssa, k, leg = reference_case()
    sc = oracle_mod.apply_delta_m_scaling(2, ssa, k, leg)      # config.delta_m_scaling = True, order = num_streams
    two = oracle_mod.twostream_radiance(**geo, ssa=sc["ssa"], ext=sc["ext"], leg=sc["leg"], f=sc["f"])["radiance"]
    do = oracle_mod.do_radiance(nstr=2, **geo, ssa=sc["ssa"], ext=sc["ext"], leg=sc["leg"], f=sc["f"],
                                include_ss=False)["radiance"]
    assert np.all(two > 0)
    np.testing.assert_allclose(two, do, rtol=2.0e-8, atol=2.0e-12)   # the reference's own tolerance
    np.testing.assert_allclose(two, do, rtol=1.0e-11)                # what the restatements achieve


@pytest.mark.parametrize("geotype", [0, 1])
def test_twostream_source_unscaled_and_pseudo_spherical(oracle_mod, geotype):
    geo, ssa, k, leg = reference_case(5)
    geo["geotype"] = geotype
    ssa = ssa * np.linspace(0.6, 1.0, ssa.shape[0])[:, None]     # altitude-dependent single-scatter albedo
    two = oracle_mod.twostream_radiance(**geo, ssa=ssa, ext=k, leg=leg)["radiance"]
    do = oracle_mod.do_radiance(nstr=2, **geo, ssa=ssa, ext=k, leg=leg, include_ss=False)["radiance"]
    np.testing.assert_allclose(two, do, rtol=1.0e-10)


def test_twostream_resonant_branches_are_continuous(oracle_mod):
    """A layer whose optical depth drives |difference x thickness| below 1e-5 switches every multiplier to the
    series forms (exp_difference / integrated_exp_difference, cpp_twostream_source.cpp:33-131): the radiance must
    stay continuous across the switch and still agree with the two-stream discrete-ordinates solve."""
    geo, ssa, k, leg = reference_case(3)
    thin = k.copy()
    thin[-3:] *= 1e-9                                        # top layers: optical depth ~1e-13
    two = oracle_mod.twostream_radiance(**geo, ssa=ssa, ext=thin, leg=leg)["radiance"]
    do = oracle_mod.do_radiance(nstr=2, **geo, ssa=ssa, ext=thin, leg=leg, include_ss=False)["radiance"]
    np.testing.assert_allclose(two, do, rtol=1.0e-9)
