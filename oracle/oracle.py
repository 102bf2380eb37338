"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes front-end of the CPU restatement in oracle/disco_oracle.hpp (built by oracle/Makefile into
oracle/liboracle_disco.so).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product (sasktran2_b200/) never does.

The eigen-solver is LAPACK dgeev from the OpenBLAS that ships inside scipy (symbol scipy_dgeev_),
located at import time and handed to the C++ side as a function pointer.
"""
from __future__ import annotations

import ctypes
import glob
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None
_DGEEV = None
_BLAS = None


def build(force: bool = False) -> Path:
    so = _HERE / "liboracle_disco.so"
    srcs = [_HERE / "disco_oracle_capi.cpp", _HERE / "disco_oracle.hpp", _HERE / "twostream_oracle.hpp", _HERE / "limb_oracle.hpp"]
    if force or not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-C", str(_HERE), "-B" if force else "-s"], check=True, capture_output=True)
    return so


def _find_openblas() -> str:
    import scipy  # noqa: F401  (only to locate scipy.libs)

    root = Path(scipy.__file__).resolve().parent.parent / "scipy.libs"
    cands = sorted(glob.glob(str(root / "libscipy_openblas*.so")))
    if not cands:
        raise RuntimeError("oracle: scipy's bundled OpenBLAS (dgeev) not found")
    return cands[0]


def lib():
    global _LIB, _DGEEV, _BLAS
    if _LIB is None:
        so = build()
        _BLAS = ctypes.CDLL(_find_openblas(), mode=ctypes.RTLD_GLOBAL)
        _DGEEV = ctypes.cast(getattr(_BLAS, "scipy_dgeev_"), ctypes.c_void_p)
        # keep OpenBLAS single-threaded: the oracle parallelises over wavelengths itself
        try:
            _BLAS.scipy_openblas_set_num_threads(1)
        except AttributeError:
            pass
        _LIB = ctypes.CDLL(str(so))
        _LIB.oracle_last_error.restype = ctypes.c_char_p
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) if a is not None else None


def do_radiance(*, nstr, alt, interp, geotype, cos_sza, earth_radius=6372000.0, los_cos_vza, los_rel_az,
                ssa, ext, leg, solar=None, albedo, d_leg=None, include_ss=True, num_azimuth=0,
                calc_derivs=False, nthreads=0, return_lanes=False, stable=False, f=None, d_f=None, reverse=False,
                brdf_kind=0, brdf_args=None, emission=None, surface_emission=None):
    """Run the oracle.

    ssa, ext: [nloc, nwavel] (Fortran order is used internally, as the reference does);
    leg: [nleg, nloc, nwavel]; d_leg: [nleg, nloc, nwavel, ngroups] or None; albedo: [nwavel].
    Returns dict(radiance [nwavel, nlos], native [nwavel, nlos, nloc*(2+G)+1] if calc_derivs).
    f [nloc, nwavel], d_f [nloc, nwavel, ngroups]: delta-M truncation fraction and its derivatives as left by
    apply_delta_m_scaling (below); the arrays passed in are then the SCALED ones.
    reverse=True computes the derivatives in reverse mode (config.do_backprop = true: layer-local duals, one transposed
    band solve per line of sight, RTESolver::backprop) instead of dense forward-mode duals; same results.
    emission [nloc, nwavel] / surface_emission [nwavel]: thermal sources of config.emission_source = DiscreteOrdinates
    (radiances only).
    stable=True switches the particular-solution multipliers from the reference's formulas to the
    singularity-free phi/psi forms (see disco_oracle.hpp, "stable multipliers"); default is the reference's.
    """
    L = lib()
    alt = np.ascontiguousarray(alt, dtype=np.float64)
    nloc = alt.size
    ssa = np.asfortranarray(ssa, dtype=np.float64)
    ext = np.asfortranarray(ext, dtype=np.float64)
    leg = np.asfortranarray(leg, dtype=np.float64)
    nwavel = ssa.shape[1]
    nleg = leg.shape[0]
    assert ssa.shape == (nloc, nwavel) and ext.shape == (nloc, nwavel) and leg.shape == (nleg, nloc, nwavel)
    cz = np.ascontiguousarray(los_cos_vza, dtype=np.float64)
    az = np.ascontiguousarray(los_rel_az, dtype=np.float64)
    nlos = cz.size
    solar = np.ones(nwavel) if solar is None else np.ascontiguousarray(solar, dtype=np.float64)
    albedo = np.ascontiguousarray(np.broadcast_to(albedo, (nwavel,)), dtype=np.float64)
    G = 0
    if d_leg is not None:
        d_leg = np.asfortranarray(d_leg, dtype=np.float64)
        G = d_leg.shape[3]
        assert d_leg.shape == (nleg, nloc, nwavel, G)
    rad = np.zeros((nwavel, nlos))
    native = lanes = None
    nl = nloc - 1
    if calc_derivs:
        native = np.zeros((nwavel, nlos, nloc * (2 + G) + 1))
        if return_lanes:
            lanes = np.zeros((nwavel, nlos, nl * (G + 2) + 1))
    L.oracle_set_stable_multipliers(ctypes.c_int(int(stable)))
    L.oracle_set_reverse_mode(ctypes.c_int(int(reverse)))
    if f is not None:
        f = np.asfortranarray(f, dtype=np.float64)
        assert f.shape == (nloc, nwavel)
        if d_f is not None:
            d_f = np.asfortranarray(d_f, dtype=np.float64)
            assert d_f.shape == (nloc, nwavel, G)
    L.oracle_set_delta_m(_p(f), _p(d_f) if f is not None else None)
    if brdf_kind:
        brdf_args = np.asfortranarray(brdf_args, dtype=np.float64)   # [nargs, nwavel]
        assert brdf_args.ndim == 2 and brdf_args.shape[1] == nwavel
        L.oracle_set_brdf(ctypes.c_int(int(brdf_kind)), ctypes.c_int(brdf_args.shape[0]), _p(brdf_args))
    if emission is not None:
        emission = np.asfortranarray(emission, dtype=np.float64)
        assert emission.shape == (nloc, nwavel)
    if surface_emission is not None:
        surface_emission = np.ascontiguousarray(np.broadcast_to(surface_emission, (nwavel,)), dtype=np.float64)
    L.oracle_set_emission(_p(emission), _p(surface_emission))
    rc = L.oracle_do_radiance(
        ctypes.c_int(nstr), ctypes.c_int(nloc), ctypes.c_int(nwavel), ctypes.c_int(nleg), ctypes.c_int(nlos),
        _p(alt), ctypes.c_int(interp), ctypes.c_int(geotype), ctypes.c_double(cos_sza), ctypes.c_double(earth_radius),
        _p(cz), _p(az), _p(ssa), _p(ext), _p(leg), _p(solar), _p(albedo), _p(d_leg), ctypes.c_int(G),
        ctypes.c_int(int(include_ss)), ctypes.c_int(num_azimuth), ctypes.c_int(int(calc_derivs)),
        ctypes.c_int(nthreads), _DGEEV, _p(rad), _p(native), _p(lanes))
    L.oracle_set_delta_m(None, None)
    L.oracle_set_brdf(ctypes.c_int(0), ctypes.c_int(1), None)
    L.oracle_set_emission(None, None)
    if rc != 0:
        raise RuntimeError(f"oracle failed ({rc}): {L.oracle_last_error().decode()}")
    out = {"radiance": rad}
    if native is not None:
        out["native"] = native
    if lanes is not None:
        out["lanes"] = lanes
    return out


def twostream_radiance(*, alt, interp, geotype, cos_sza, earth_radius=6372000.0, los_cos_vza, los_rel_az, ssa, ext, leg,
                       solar=None, albedo, f=None, nthreads=0, **_ignored):
    """The reference's dedicated two-stream source (multiple scatter only; twostream_oracle.hpp).  Returns
    dict(radiance [nwavel, nlos])."""
    L = lib()
    alt = np.ascontiguousarray(alt, dtype=np.float64)
    ssa = np.asfortranarray(ssa, dtype=np.float64)
    ext = np.asfortranarray(ext, dtype=np.float64)
    leg = np.asfortranarray(leg, dtype=np.float64)
    cz = np.ascontiguousarray(los_cos_vza, dtype=np.float64)
    az = np.ascontiguousarray(los_rel_az, dtype=np.float64)
    nloc, nwavel = ssa.shape
    solar = np.ones(nwavel) if solar is None else np.ascontiguousarray(solar, dtype=np.float64)
    albedo = np.ascontiguousarray(np.broadcast_to(albedo, (nwavel,)), dtype=np.float64)
    rad = np.zeros((nwavel, cz.size))
    if f is not None:
        f = np.asfortranarray(f, dtype=np.float64)
    L.oracle_set_delta_m(_p(f), None)
    rc = L.oracle_twostream_radiance(ctypes.c_int(nloc), ctypes.c_int(nwavel), ctypes.c_int(leg.shape[0]), ctypes.c_int(cz.size),
                                     _p(alt), ctypes.c_int(interp), ctypes.c_int(geotype), ctypes.c_double(cos_sza),
                                     ctypes.c_double(earth_radius), _p(cz), _p(az), _p(ssa), _p(ext), _p(leg), _p(solar),
                                     _p(albedo), ctypes.c_int(nthreads), _p(rad))
    L.oracle_set_delta_m(None, None)
    if rc != 0:
        raise RuntimeError(f"oracle_twostream_radiance failed: {L.oracle_last_error().decode()}")
    return {"radiance": rad}


def degeneracy(*, nstr, alt, interp, geotype, cos_sza, earth_radius=6372000.0, los_cos_vza, los_rel_az, ssa, ext, leg,
               f=None, **_ignored):
    """Distances [nwavel, nstr (order), L (layer), 2] of every cell to the removable singularities of the reference's
    multiplier formulas: [..., 0] = min_j |secant - k_j|, [..., 1] = min_{j, los} |1 - mu_los k_j|."""
    L = lib()
    alt = np.ascontiguousarray(alt, dtype=np.float64)
    ssa = np.asfortranarray(ssa, dtype=np.float64)
    ext = np.asfortranarray(ext, dtype=np.float64)
    leg = np.asfortranarray(leg, dtype=np.float64)
    cz = np.ascontiguousarray(los_cos_vza, dtype=np.float64)
    az = np.ascontiguousarray(los_rel_az, dtype=np.float64)
    nloc, nwavel = ssa.shape
    out = np.zeros((nwavel, nstr, nloc - 1, 2))
    if f is not None:
        f = np.asfortranarray(f, dtype=np.float64)
    L.oracle_set_delta_m(_p(f), None)
    rc = L.oracle_degeneracy(ctypes.c_int(nstr), ctypes.c_int(nloc), ctypes.c_int(nwavel), ctypes.c_int(leg.shape[0]),
                             ctypes.c_int(cz.size), _p(alt), ctypes.c_int(interp), ctypes.c_int(geotype),
                             ctypes.c_double(cos_sza), ctypes.c_double(earth_radius), _p(cz), _p(az), _p(ssa), _p(ext),
                             _p(leg), _DGEEV, _p(out))
    L.oracle_set_delta_m(None, None)
    if rc != 0:
        raise RuntimeError(f"oracle_degeneracy failed: {L.oracle_last_error().decode()}")
    return out


def plan(*, nstr, alt, interp, geotype, cos_sza, earth_radius=6372000.0, los_cos_vza, los_rel_az,
         chapman_straight_line=False):
    """Geometry plan of the oracle.  Pseudo-spherical chapman factors are ray traced like the reference
    (calculate_chapman_factors_raytracer); chapman_straight_line=True selects the closed formula instead."""
    L = lib()
    L.oracle_set_chapman_straight_line(ctypes.c_int(int(chapman_straight_line)))
    alt = np.ascontiguousarray(alt, dtype=np.float64)
    nloc = alt.size
    cz = np.ascontiguousarray(los_cos_vza, dtype=np.float64)
    az = np.ascontiguousarray(los_rel_az, dtype=np.float64)
    nlos = cz.size
    N = nstr // 2
    nl = nloc - 1
    out = dict(mu=np.zeros(nstr), wt=np.zeros(nstr), lp_mu=np.zeros((nstr, N, nstr)), lp_csz=np.zeros((nstr, nstr)),
               lp_los=np.zeros((nlos, nstr, nstr)), W=np.zeros((nl, nloc)), chapman=np.zeros((nl, nl)))
    rc = L.oracle_plan(ctypes.c_int(nstr), ctypes.c_int(nloc), ctypes.c_int(nlos), _p(alt), ctypes.c_int(interp),
                       ctypes.c_int(geotype), ctypes.c_double(cos_sza), ctypes.c_double(earth_radius), _p(cz), _p(az),
                       _p(out["mu"]), _p(out["wt"]), _p(out["lp_mu"]), _p(out["lp_csz"]), _p(out["lp_los"]),
                       _p(out["W"]), _p(out["chapman"]))
    L.oracle_set_chapman_straight_line(ctypes.c_int(0))
    if rc != 0:
        raise RuntimeError(f"oracle_plan failed: {L.oracle_last_error().decode()}")
    return out


def band_solve(a_dense, b, kl, trans=False):
    L = lib()
    a = np.ascontiguousarray(a_dense, dtype=np.float64)
    x = np.array(b, dtype=np.float64, copy=True)
    rc = L.oracle_band_solve(ctypes.c_int(a.shape[0]), ctypes.c_int(kl), _p(a), _p(x), ctypes.c_int(int(trans)))
    if rc != 0:
        raise RuntimeError(f"band solve info={rc}")
    return x


def apply_mappings(native, mappings, nloc, ngroups):
    """native [nwavel, nlos, nnative] -> dict name -> WF [nout, nwavel, nlos] following OutputC::assign_lane
    (cpp/lib/output/outputc.cpp:37-160).  mappings: dict name -> dict(d_ssa [nloc,nw], d_extinction [nloc,nw],
    scat_factor [nloc,nw] or None, scat_index int, interpolator [nloc, nout] or None).
    """
    out = {}
    d_k = native[:, :, 0:nloc]
    d_w = native[:, :, nloc:2 * nloc]
    for name, mp in mappings.items():
        s = d_w * mp["d_ssa"].T[:, None, :] + d_k * mp["d_extinction"].T[:, None, :]
        if mp.get("scat_factor") is not None:
            g = mp["scat_index"]
            d_s = native[:, :, 2 * nloc + g * nloc: 2 * nloc + (g + 1) * nloc]
            s = s + d_s * mp["scat_factor"].T[:, None, :]
        if mp.get("interpolator") is not None:
            s = s @ mp["interpolator"]
        out[name] = np.ascontiguousarray(np.moveaxis(s, 2, 0))
    return out


def apply_delta_m_scaling(order, ssa, ext, leg, d_leg=None, mappings=None):
    """numpy restatement of Atmosphere::apply_delta_m_scaling (cpp/lib/atmosphere/atmosphere.cpp:69-203).

    ssa, ext: [nloc, nwavel]; leg: [nleg, nloc, nwavel]; d_leg: [nleg, nloc, nwavel, G] or None; mappings: dict
    name -> dict(d_ssa, d_extinction, scat_factor or None, scat_index).  Returns scaled COPIES:
    dict(ssa, ext, leg, d_leg, f, d_f, mappings); the inputs are left untouched.  order >= nleg: unscaled (f None)."""
    ssa0 = np.array(ssa, dtype=np.float64)
    ext0 = np.array(ext, dtype=np.float64)
    leg = np.array(leg, dtype=np.float64)
    d_leg = None if d_leg is None else np.array(d_leg, dtype=np.float64)
    maps = {k: {kk: (np.array(vv, dtype=np.float64) if isinstance(vv, np.ndarray) else vv) for kk, vv in v.items()}
            for k, v in (mappings or {}).items()}
    if order >= leg.shape[0]:
        return dict(ssa=ssa0, ext=ext0, leg=leg, d_leg=d_leg, f=None, d_f=None, mappings=maps)
    f = leg[order] / (2 * order + 1)                                  # :96-103
    ext1 = ext0 * (1 - ssa0 * f)                                      # :106-109
    ssa1 = (1 - f) / (1 - ssa0 * f) * ssa0                            # :112-116
    d_f = None
    if d_leg is not None:
        d_f = d_leg[order] / (2 * order + 1)                          # :119-128  [nloc, nwavel, G]
    leg = leg / (1 - f)[None]                                         # :140
    if d_leg is not None:
        d_leg = (d_leg + leg[..., None] * d_f[None]) / (1 - f)[None, :, :, None]   # :147-151
    for m in maps.values():                                           # :158-201
        if m.get("d_extinction") is None:
            continue
        dk = m["d_extinction"] * (1 - ssa0 * f)
        dk = dk - ext0 * f * m["d_ssa"]
        dw = m["d_ssa"] * (1 - f * (1 - ssa1)) / (1 - ssa0 * f)
        if m.get("scat_factor") is not None and m.get("scat_index", -1) >= 0:
            df = d_f[:, :, m["scat_index"]] * m["scat_factor"]
            dk = dk - ssa0 * ext0 * df
            dw = dw + df * ssa0 / (1 - ssa0 * f) * (ssa1 - 1)
        m["d_extinction"], m["d_ssa"] = dk, dw
    return dict(ssa=ssa1, ext=ext1, leg=leg, d_leg=d_leg, f=f, d_f=d_f, mappings=maps)


def _ray_table(rays):
    """rays: list of ("ground", cos_sza, rel_az, cos_vza, observer_altitude) / ("tangent", tangent_altitude, rel_az,
    observer_altitude, cos_sza) in the argument order of the reference's GroundViewingSolar / TangentAltitudeSolar."""
    tab = np.zeros((len(rays), 5))
    for i, r in enumerate(rays):
        tab[i, 0] = {"ground": 0.0, "tangent": 1.0}[r[0]]
        tab[i, 1:] = r[1:5]
    return tab


def limb_radiance(*, nstr, alt, interp, cos_sza, saa=0.0, earth_radius=6372000.0, rays, num_sza=2, ms_do=True,
                  ss_exact=False, num_ss_moments=16, ssa, ext, leg, solar=None, albedo, nthreads=0, exact_tangent=False):
    """Spherical line-of-sight path (oracle/limb_oracle.hpp): DO multiple-scatter source table interpolated onto the
    traced rays (+ exact single scatter when ss_exact).  Returns dict(radiance [nwavel, nrays], los_optical_depth).
    exact_tangent=True removes the reference's rounding-dependent ~0.1 m error of the tangent-layer lengths (see
    limb_oracle.hpp, exact_tangent_ref); default is the reference's arithmetic."""
    L = lib()
    L.oracle_set_exact_tangent(int(exact_tangent))
    alt = np.ascontiguousarray(alt, dtype=np.float64)
    ssa = np.asfortranarray(ssa, dtype=np.float64)
    ext = np.asfortranarray(ext, dtype=np.float64)
    leg = np.asfortranarray(leg, dtype=np.float64)
    nloc, nwavel = ssa.shape
    nleg = leg.shape[0]
    albedo = np.ascontiguousarray(albedo, dtype=np.float64)
    solar = np.ones(nwavel) if solar is None else np.ascontiguousarray(solar, dtype=np.float64)
    tab = _ray_table(rays)
    rad = np.zeros((nwavel, len(rays)))
    od = np.zeros((nwavel, len(rays)))
    rc = L.oracle_limb_radiance(nstr, nloc, nwavel, nleg, len(rays), _p(alt), int(interp), ctypes.c_double(cos_sza),
                                ctypes.c_double(saa), ctypes.c_double(earth_radius), _p(tab), int(num_sza), int(ms_do),
                                int(ss_exact), int(num_ss_moments), _p(ssa), _p(ext), _p(leg), _p(solar), _p(albedo), _DGEEV,
                                int(nthreads), _p(rad), _p(od))
    if rc != 0:
        raise RuntimeError(f"oracle_limb_radiance failed: {L.oracle_last_error().decode()}")
    return dict(radiance=rad, los_optical_depth=od)


def limb_geometry(*, alt, interp, cos_sza, saa=0.0, earth_radius=6372000.0, rays, max_layers=512, exact_tangent=False):
    """Traced-ray geometry of the limb oracle: dict(nlayers [nrays], ground_hit [nrays], layers [nrays, max_layers, 9])
    with per layer (layer_distance, od_quad_start, od_quad_end, cos_sza_entrance, cos_sza_exit, saz_entrance, saz_exit,
    r_entrance, r_exit); layers[0] is the one farthest from the observer."""
    L = lib()
    L.oracle_set_exact_tangent(int(exact_tangent))
    alt = np.ascontiguousarray(alt, dtype=np.float64)
    tab = _ray_table(rays)
    n = len(rays)
    nl = np.zeros(n, dtype=np.int32)
    gh = np.zeros(n, dtype=np.int32)
    data = np.zeros((n, max_layers, 9))
    csc = np.zeros(n)
    ip = ctypes.POINTER(ctypes.c_int)
    rc = L.oracle_limb_geometry(alt.size, n, _p(alt), int(interp), ctypes.c_double(cos_sza), ctypes.c_double(saa),
                                ctypes.c_double(earth_radius), _p(tab), int(max_layers), nl.ctypes.data_as(ip),
                                gh.ctypes.data_as(ip), _p(data), _p(csc))
    if rc != 0:
        raise RuntimeError(f"oracle_limb_geometry failed: {L.oracle_last_error().decode()}")
    return dict(nlayers=nl, ground_hit=gh, layers=data, cos_scatter=csc)


def brdf_value(kind, args, mu_in, mu_out, phi_diff):
    """BRDF model of the reference (cpp/include/sasktran2/atmosphere/surface.h): kind 0 Lambertian, 1 snow, 2 MODIS."""
    L = lib()
    L.oracle_brdf_value.restype = ctypes.c_double
    a = np.ascontiguousarray(args, dtype=np.float64)
    return L.oracle_brdf_value(int(kind), _p(a), ctypes.c_double(mu_in), ctypes.c_double(mu_out), ctypes.c_double(phi_diff))


def brdf_expansion(m, kind, args, mu_out, mu_in):
    """Azimuthal Fourier coefficient rho_m(mu_out, mu_in) (SurfaceStorage::compute_expansion, 512-point quadrature)."""
    L = lib()
    L.oracle_brdf_expansion.restype = ctypes.c_double
    a = np.ascontiguousarray(args, dtype=np.float64)
    return L.oracle_brdf_expansion(int(m), int(kind), _p(a), ctypes.c_double(mu_out), ctypes.c_double(mu_in))
