#!/usr/bin/env python
"""Benchmark of the B200 discrete-ordinates radiance solve (contract: see the task statement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c5|c2|c1|c3]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one batch of synthetic input.  Default workload = BASELINE.json configs[4]
("C5", the configuration the metric is quoted on): full linearisation, pseudo-spherical DO, 16 streams, 100 layers, 10
ground-viewing LOS, weighting functions w.r.t. O3 VMR, NO2 VMR and aerosol extinction (3 x 101 outputs) + albedo,
50 000 wavelengths.  With N GPUs the SAME spectrum is sharded into contiguous wavelength blocks, one process per GPU
(strong scaling); the only exchange is the final gather of radiances and weighting functions onto rank 0.

  value  LOS x wavelength radiances / s, inputs resident in HBM, kernels only (CUDA events inside the library, max
         over ranks)
  e2e    the same metric through the reference-facing call (sk_engine_calculate_radiance / ..._block_thread behind
         Engine.calculate_radiance) with HOST buffers: H2D of the inputs, the solve, and the results of ALL ranks ending
         up in rank 0's host arrays, inside the timed region.  N > 1 reports both gather variants: "nccl" (ncclSend /
         ncclRecv onto GPU 0 inside the library, then one D2H) and "direct" (every rank copies its block straight into
         the caller's page-locked shared result arrays: N PCIe links instead of one); e2e.value is the faster one.
  --impl reference   the CPU implementation of the path on the SAME config: the oracle port of the reference's
         reverse-mode linearisation (RTESolver::backprop), OpenMP over wavelengths on all host cores, on a bounded
         sample of the workload (the reference itself cannot be built here: Eigen + Rust/cxx generated headers).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = json.loads((ROOT / "BASELINE.json").read_text())["metric"] if (ROOT / "BASELINE.json").exists() else \
    "LOS×wavelength radiances/sec (16-stream, 100 layers, with WFs) at 1/2/4/8 B200"
UNIT = "radiances/s"

# Named workloads (BASELINE.json configs; SURVEY.md section 8d fixes the unspecified details)
CONFIGS = {
    "c5": dict(label="configs[4] (C5)", nwavel=50000, nstr=16, layers=100, nlos=10, wf=True,
               text="full linearisation: pseudo-spherical DO, 16 streams, 100 layers, Rayleigh+aerosol+O3/NO2, 10 ground-viewing "
                    "LOS, weighting functions w.r.t. O3 VMR, NO2 VMR, aerosol extinction (3 x 101 outputs) and albedo, 50,000 "
                    "wavelengths"),
    "c2": dict(label="configs[1] (C2)", nwavel=100000, nstr=16, layers=100, nlos=10, wf=False,
               text="pseudo-spherical DO, 16 streams, 100 layers, Rayleigh+aerosol, 10 nadir LOS, 100,000 wavelengths, radiances only"),
    "c1": dict(label="configs[0] (C1)", nwavel=1000, nstr=4, layers=50, nlos=1, wf=False,
               text="plane-parallel DO, 4 streams, 50 layers, Rayleigh+O3, 1 nadir LOS, 1,000 wavelengths"),
    "c4": dict(label="configs[3] (C4)", nwavel=10000, nstr=16, layers=100, nlos=100, wf=False,
               text="OSIRIS-style limb: spherical geometry, exact single scatter + 16-stream DO multiple-scatter source table "
                    "(2 SZAs), 100 tangent-altitude LOS (10-60 km), 100 layers, 10,000 wavelengths"),
    "c3": dict(label="configs[2] (C3)", nwavel=1000000, nstr=2, layers=60, nlos=2, wf=False,
               text="two-stream source (num_streams=2, multiple scatter only), 60 layers, 2 nadir LOS, 1,000,000 line-by-line "
                    "wavelengths (O2 A-band like)"),
}


def flop_model(nstr, nlayers, nlos, m_list, ngroups=1, adjoint_refactor=False):
    """Algorithmic flops per wavelength, split by timed kernel group.

    layer, bvp: SURVEY.md section 8d / Appendix B (values only; the BVP figure is LAPACK's banded LU + solve count,
    the staircase elimination does about a third of it).  Weighting functions (reverse mode):
      wf_adjoint  one banded solve (LAPACK dgbtrs count) per line of sight with the forward factors, as the
                  reference's backprop does; `adjoint_refactor` adds the second factorisation (of A^T) that the
                  SK_B200_ADJOINT=refactor path and few-LOS shapes perform;
      wf_layer    the layer-local linearisation with NL = ngroups + 4 lanes [eps_g | tau | omega | t | s]: NL times the
                  values-only particular + LOS work (forward-mode count of the reference's layer duals,
                  sktran_do_rte.cpp:903-1332, sktran_do_opticallayer.cpp:94-555) + 7 N^3 per eigen-derivative lane
                  (ngroups + 1 lanes, sktran_do_rte.cpp:198-298) + the adjoint contraction 8 N^2 NL per LOS."""
    N, K, L = nstr // 2, nstr, nlayers
    NL, NH = ngroups + 4, ngroups + 1
    layer = bvp = wf_adjoint = wf_layer = 0.0
    for m in m_list:
        homog = L * (6 * N * N * (K - m) + 29 * N**3 + 4 * N * N)            # S+-, eigen-decomposition, W+-
        part = L * (12 * N * (K - m) + 10 * N * N + 60 * N)                  # Green's function particular solution
        post = nlos * L * (6 * N * (K - m) + 8 * N * N + 120 * N)            # LOS source multipliers
        factor = 4 * N * L * (3 * N - 1) * (6 * N - 2)                       # banded LU (LAPACK count)
        solve = 4 * N * L * (9 * N - 3)                                      # one banded solve
        layer += homog + part + post
        bvp += factor + solve
        wf_adjoint += (factor if adjoint_refactor else 0) + nlos * solve
        wf_layer += NL * (part + post) + L * NH * 7 * N**3 + nlos * L * 8 * N * N * NL
    return {"layer": layer, "bvp": bvp, "wf_adjoint": wf_adjoint, "wf_layer": wf_layer, "total": layer + bvp,
            "total_wf": layer + bvp + wf_adjoint + wf_layer}


# DRAM bytes per wavelength (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture, divided by
# the wavelengths of the profiled launch) for the C5 shape - see NCU_SOURCE.  Other shapes report traffic = null.
NCU_SOURCE = "profiles/ncu_r01_v12_summary.csv"
NCU_DRAM_BYTES_PER_WAVELENGTH = {"layer": (0.813 + 2.020 + 3.324) * 1e9 / 600, "bvp": (6.151 + 7.813) * 1e9 / 600,
                                 "wf_adjoint": (10.033 + 2.442) * 1e9 / 600, "wf_layer": (3.041 + 0.453) * 1e9 / 600}
try:  # a newer capture of this round, written by tools/ncu_summary.py
    _t = json.loads((ROOT / "profiles" / "ncu_dram_bytes_per_wavelength.json").read_text())
    NCU_DRAM_BYTES_PER_WAVELENGTH, NCU_SOURCE = _t["bytes_per_wavelength"], _t["source"]
except Exception:
    pass


def tsolve_bytes_model(nstr, nlayers, nlos, n_orders):
    """Algorithmic HBM bytes per wavelength of the transposed solves (k_bvp_tsolve): per
    (order, pivot) one factor row (FS doubles) and one multiplier row (LS doubles) read once, per (order, LOS) the
    right-hand side read, y written and read back, z written (2 N L doubles each)."""
    N = nstr // 2
    fs = ((4 * N + 2) & ~1) + 2
    ls = (3 * N + 2) & ~1
    n = 2 * N * nlayers
    return 8.0 * n_orders * (n * (fs + ls) + nlos * 4 * n)


def bytes_model(nloc, nleg, nlos, nwf_out=0, ngroups=0):
    """Algorithmic HBM bytes per wavelength: inputs 8 nloc (2 + nleg)(1 + ngroups) + outputs 8 nlos (1 + sum nout)."""
    return 8.0 * nloc * (2 + nleg) * (1 + ngroups) + 8.0 * nlos * (1 + nwf_out)


class ClockSampler:
    """Samples SM clocks / throttle reasons of one GPU during the timed region (pynvml)."""

    def __init__(self, index):
        self.samples, self.reasons, self.maxclk = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.maxclk = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        if self.nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.maxclk, "reasons": sorted(self.reasons)}


# ---------------------------------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------------------------------
def build_scenario(cfg_name, nw_total, block, nlos=None, layers=None, nstr=None):
    from sasktran2_b200 import scenarios

    c = CONFIGS[cfg_name]
    nlos = c["nlos"] if nlos is None else nlos
    layers = c["layers"] if layers is None else layers
    nstr = c["nstr"] if nstr is None else nstr
    if cfg_name in ("c5", "c2"):
        return scenarios.config2(nwavel=nw_total, nlayers=layers, nstr=nstr, nlos=nlos, with_wf=c["wf"], block=block)
    if cfg_name == "c1":
        sc = scenarios.config1(nwavel=nw_total, nlayers=layers)
        assert block is None or block == (0, nw_total), "config 1 is a single-GPU case"
        return sc
    if cfg_name == "c3":
        return scenarios.config3(nwavel=nw_total, nlayers=layers, nlos=nlos, block=block)
    if cfg_name == "c4":
        return scenarios.config4(nwavel=nw_total, nlayers=layers, nstr=nstr, nrays=nlos, block=block)
    raise ValueError(cfg_name)


def make_engine(sc, cfg_name, total_wavelengths=None, start=0):
    """(config, geometry, viewing geometry, engine, atmosphere) for a scenario block of one named workload."""
    import sasktran2_b200 as sk

    cfg = sk.Config()
    cfg.num_streams = sc.nstr
    cfg.num_stokes = 1
    if cfg_name == "c3":
        cfg.multiple_scatter_source = sk.MultipleScatterSource.TwoStream
        cfg.single_scatter_source = sk.SingleScatterSource.NoSource
    elif cfg_name == "c4":
        cfg.multiple_scatter_source = sk.MultipleScatterSource.DiscreteOrdinates
        cfg.single_scatter_source = sk.SingleScatterSource.Exact
        cfg.num_sza = sc.num_sza
        cfg.num_singlescatter_moments = sc.leg_coeff.shape[0]
    else:
        cfg.multiple_scatter_source = sk.MultipleScatterSource.DiscreteOrdinates
        cfg.single_scatter_source = sk.SingleScatterSource.DiscreteOrdinates
    cfg.do_backprop = True
    geo = sk.Geometry1D(sc.cos_sza, sc.saa, sc.earth_radius, sc.altitudes, sk.InterpolationMethod(sc.interp),
                        sk.GeometryType(sc.geotype))
    view = sk.ViewingGeometry()
    for cz, az in zip(sc.los_cos_vza, sc.los_rel_az):
        view.add_ray(sk.GroundViewingSolar(sc.cos_sza, float(az), float(cz), sc.observer_altitude))
    for r in sc.rays:
        view.add_ray(sk.TangentAltitudeSolar(*r[1:5]) if r[0] == "tangent" else sk.GroundViewingSolar(*r[1:5]))
    eng = sk.Engine(cfg, geo, view)
    atm = sk.Atmosphere.from_scenario(sc, geo, cfg, total_wavelengths=total_wavelengths, wavelength_start=start)
    if sc.mappings:
        atm.surface.enable_albedo_derivative("wf_albedo")
    return cfg, geo, view, eng, atm


def oracle_inputs(sc, pick):
    return dict(nstr=sc.nstr, alt=sc.altitudes, interp=sc.interp, geotype=sc.geotype, cos_sza=sc.cos_sza,
                earth_radius=sc.earth_radius, los_cos_vza=sc.los_cos_vza, los_rel_az=sc.los_rel_az,
                ssa=np.asfortranarray(sc.ssa[:, pick]), ext=np.asfortranarray(sc.total_extinction[:, pick]),
                leg=np.asfortranarray(sc.leg_coeff[:, :, pick]), albedo=sc.albedo[pick])


def run_oracle(sc, pick, threads, cfg_name, stable=False):
    """The CPU port on wavelengths `pick` of the scenario: same outputs as the GPU arm of that workload (radiances;
    for C5 also the native derivatives in reverse mode, do_backprop = true).  Returns (result dict, seconds)."""
    from oracle import oracle

    inp = oracle_inputs(sc, pick)
    extra = {}
    if cfg_name == "c3":
        t0 = time.perf_counter()
        out = oracle.twostream_radiance(**{k: v for k, v in inp.items() if k != "nstr"}, nthreads=threads)
        return out, time.perf_counter() - t0
    if cfg_name == "c4":
        t0 = time.perf_counter()
        out = oracle.limb_radiance(nstr=sc.nstr, alt=sc.altitudes, interp=sc.interp, cos_sza=sc.cos_sza, saa=sc.saa,
                                   earth_radius=sc.earth_radius, rays=sc.rays, num_sza=sc.num_sza, ms_do=True, ss_exact=True,
                                   num_ss_moments=sc.leg_coeff.shape[0], ssa=inp["ssa"], ext=inp["ext"], leg=inp["leg"],
                                   albedo=inp["albedo"], nthreads=threads, exact_tangent=stable)
        return out, time.perf_counter() - t0
    if CONFIGS[cfg_name]["wf"]:
        aer = sc.mappings["wf_aerosol_extinction"]["d_legendre"]
        extra = dict(d_leg=np.asfortranarray(aer[:, :, pick])[..., None], calc_derivs=True, reverse=True, stable=stable)
    t0 = time.perf_counter()
    out = oracle.do_radiance(**inp, nthreads=threads, **extra)
    return out, time.perf_counter() - t0


def oracle_weighting_functions(sc, pick, native):
    from oracle import oracle

    maps = {}
    for n, mp in sc.mappings.items():
        maps[n] = dict(d_ssa=np.asfortranarray(mp["d_ssa"][:, pick]), d_extinction=np.asfortranarray(mp["d_extinction"][:, pick]),
                       scat_factor=np.asfortranarray(mp["scat_factor"][:, pick]) if "scat_factor" in mp else None,
                       scat_index=0 if "scat_factor" in mp else -1, interpolator=None)
    wf = oracle.apply_mappings(native, maps, sc.nloc, 1)
    wf["wf_albedo"] = native[:, :, -1]
    return wf


def parity_sample(sc, cfg_name, result, nsample, threads):
    """Compares a strided sample of the TIMED run's outputs (this rank's block) with the oracle: radiance 1e-9; for
    C5 the O3 / NO2 VMR and albedo weighting functions 1e-7 of the column maximum (singularity-free oracle variant;
    tests/wf_checks.py holds the full rule set incl. the reference-formula variant)."""
    pick = np.unique(np.linspace(0, sc.nwavel - 1, nsample).astype(int))
    ora, _ = run_oracle(sc, pick, threads, cfg_name, stable=True)
    rad = result["radiance"][..., 0][pick]
    out = {"wavelengths": int(pick.size), "radiance_max_rel_diff": float(np.max(np.abs(rad / ora["radiance"] - 1.0))),
           "radiance_tol": 1e-9}
    ok = out["radiance_max_rel_diff"] < 1e-9
    if CONFIGS[cfg_name]["wf"]:
        wf = oracle_weighting_functions(sc, pick, ora["native"])
        errs = {}
        for name in ("wf_o3_vmr", "wf_no2_vmr", "wf_aerosol_extinction", "wf_albedo"):
            got = result[name][..., 0][:, pick] if name != "wf_albedo" else result[name][..., 0][pick]
            ref = wf[name]
            errs[name] = float(np.max(np.abs(got - ref) / np.abs(ref).max(axis=0, keepdims=True)))
        out["wf_max_diff_over_column_max"] = errs
        out["wf_tol"] = 1e-7
        out["wf_note"] = "aerosol extinction: 1/k-amplified mapping, reported only (rules: tests/wf_checks.py)"
        ok = ok and all(errs[n] < 1e-7 for n in ("wf_o3_vmr", "wf_no2_vmr", "wf_albedo"))
    out["ok"] = bool(ok)
    return out


def cpu_arm(sc, cfg_name, sample, threads):
    """Times the CPU port on a bounded strided sample of the workload; returns (radiances/s, seconds, description)."""
    pick = np.unique(np.linspace(0, sc.nwavel - 1, sample).astype(int))
    _, dt = run_oracle(sc, pick, threads, cfg_name)
    what = ("oracle port of the reference's reverse-mode linearisation (RTESolver::backprop, do_backprop = true): radiances + "
            "native derivatives w.r.t. extinction, SSA, one scattering group and albedo at every grid point"
            if CONFIGS[cfg_name]["wf"] else "oracle port, radiances only")
    return pick.size * sc.nlos / dt, dt, f"{pick.size} wavelengths x {sc.nlos} LOS strided over the workload's spectrum, {what}, " \
                                        f"OpenMP over wavelengths on {threads} threads, {dt:.1f} s"


def run_reference(args):
    """--impl reference: the CPU implementation of the path on the same config (kind "port": the reference itself needs
    Eigen + Rust/cxx generated headers and cannot be built in this image)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c = CONFIGS[args.config]
    cores = os.cpu_count() or 1
    sample = args.ref_sample
    # the sample is drawn from the first `span` wavelengths... no: build a decimated spectrum with the same formulae
    sc = build_scenario(args.config, c["nwavel"], None) if c["nwavel"] <= 4 * sample else \
        build_scenario_decimated(args.config, c["nwavel"], sample)
    cpu_arm(sc, args.config, min(sample, 2 * cores), cores)  # warm-up (thread pool, page-in)
    vals, dts = [], []
    for _ in range(max(args.steps, 1)):
        v, dt, desc = cpu_arm(sc, args.config, sample, cores)
        vals.append(v)
        dts.append(dt)
    v = float(np.median(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.median(dts)) * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{c['label']}: {c['text']}"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def build_scenario_decimated(cfg_name, nw_total, sample):
    """`sample` wavelengths strided over the nw_total-point spectrum of a workload, built block by block so that the CPU
    arms never allocate the full spectrum."""
    import copy

    pick = np.unique(np.linspace(0, nw_total - 1, sample).astype(int))
    parts = [build_scenario(cfg_name, nw_total, (int(w), 1)) for w in pick]
    sc = copy.copy(parts[0])
    sc.ssa = np.asfortranarray(np.concatenate([p.ssa for p in parts], axis=1))
    sc.total_extinction = np.asfortranarray(np.concatenate([p.total_extinction for p in parts], axis=1))
    sc.leg_coeff = np.asfortranarray(np.concatenate([p.leg_coeff for p in parts], axis=2))
    sc.albedo = np.concatenate([p.albedo for p in parts])
    sc.solar_irradiance = np.concatenate([p.solar_irradiance for p in parts])
    sc.mappings = {}
    for n in parts[0].mappings:
        sc.mappings[n] = {k: np.asfortranarray(np.concatenate([p.mappings[n][k] for p in parts], axis=-1))
                          for k in parts[0].mappings[n]}
    return sc


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def quick_config(cfg_name, steps, warmup, local_rank, threads):
    """Single-GPU line of another named workload (device-resident rate, end-to-end rate, parity sample, roofline view)."""
    import sasktran2_b200 as sk  # noqa: F401

    c = CONFIGS[cfg_name]
    t0 = time.perf_counter()
    sc = build_scenario(cfg_name, c["nwavel"], None)
    _, _, _, eng, atm = make_engine(sc, cfg_name)
    eng.set_workspace_gb(48.0)
    eng.reuse_output_buffers = True
    eng.stage(atm)
    for _ in range(warmup):
        eng.solve_staged()
    ms = 0.0
    for _ in range(steps):
        eng.solve_staged()
        ms += eng.timings_ms()["kernels_total"]
    ms /= steps
    res = eng.fetch()
    # end to end: page-locked caller buffers (sk_b200_host_register on the input arrays, result arrays from sk_b200_host_alloc)
    lib = sk._lib.lib()
    ins = [atm.storage.ssa, atm.storage.total_extinction, atm.storage.leg_coeff, atm.storage.solar_irradiance, atm.surface.albedo]
    registered = [a for a in ins if lib.sk_b200_host_register(a.ctypes.data, a.nbytes) == 0]
    eng.calculate_radiance(atm)
    t1 = time.perf_counter()
    for _ in range(steps):
        res = eng.calculate_radiance(atm)
    e2e_ms = (time.perf_counter() - t1) * 1e3 / steps
    for a in registered:
        lib.sk_b200_host_unregister(a.ctypes.data)
    units = float(sc.nwavel * sc.nlos)
    nleg = sc.leg_coeff.shape[0]
    gbs = bytes_model(sc.nloc, nleg, sc.nlos) * sc.nwavel / (ms * 1e-3) / 1e9
    out = {"workload": f"{c['label']}: {c['text']}", "value": units / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
           "e2e": {"value": units / (e2e_ms * 1e-3), "ms_per_step": e2e_ms, "pinned": len(registered) == len(ins),
                   "h2d_bytes_per_step": int(sum(a.nbytes for a in ins)), "d2h_bytes_per_step": int(8 * units)},
           "gpu_launches": int(eng.kernel_launches()),
           "hbm_view": {"achieved_gbs": gbs, "bytes_per_wavelength": bytes_model(sc.nloc, nleg, sc.nlos)},
           "kernel_ms": {k: v for k, v in eng.timings_ms().items() if v > 0 and k not in ("h2d", "d2h")},
           "parity": parity_sample(sc, cfg_name, res, 8 if cfg_name == "c4" else 16, threads), "setup_s": time.perf_counter() - t0}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default=os.environ.get("SK_BENCH_CONFIG", "c5"), choices=sorted(CONFIGS))
    ap.add_argument("--nwavel", type=int, default=int(os.environ.get("SK_BENCH_NWAVEL", "0")),
                    help="total wavelengths of the spectrum (default: the named configuration's)")
    ap.add_argument("--cpu-sample", type=int, default=2000, help="wavelengths of the cpu_baseline sample (same config; about 15 s on 16 cores)")
    ap.add_argument("--ref-sample", type=int, default=640, help="wavelengths per step of --impl reference (about 5 s on 16 cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--workspace-gb", type=float, default=48.0,
                    help="device workspace per wavelength chunk (B200: 180 GB HBM3e)")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    import sasktran2_b200 as sk
    from sasktran2_b200.parallel import SharedResult, init_nccl_comm, wavelength_block

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(x, op):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    reduce_max = lambda x: reduce(x, dist.ReduceOp.MAX if world > 1 else None)  # noqa: E731
    reduce_sum = lambda x: reduce(x, dist.ReduceOp.SUM if world > 1 else None)  # noqa: E731

    c = CONFIGS[args.config]
    if args.config in ("c1", "c3", "c4"):
        # named workloads other than the headline: one single-GPU line in the contract's format
        if world > 1:
            raise SystemExit("bench.py: --config c1 / c3 / c4 are single-GPU lines (the scaling run is the default config)")
        sk._lib.check(sk._lib.lib().sk_b200_set_device(local_rank), "set_device")
        if args.nwavel > 0:
            CONFIGS[args.config] = dict(c, nwavel=args.nwavel)
        sampler = ClockSampler(local_rank)
        sampler.start()
        q = quick_config(args.config, args.steps, args.warmup, local_rank, os.cpu_count() or 1)
        clocks = sampler.stop()
        if not q["parity"]["ok"]:
            raise SystemExit(f"bench.py: timed outputs disagree with the oracle: {q['parity']}")
        line = {"metric": METRIC, "value": q["value"], "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": q["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": {"workload": q["workload"]}, "clocks": clocks,
                "e2e": dict(q["e2e"], unit=UNIT), "gpu_launches": q["gpu_launches"], "kernel_ms_per_step": q["kernel_ms"],
                "roofline": {"bound": "hbm", "achieved": q["hbm_view"]["achieved_gbs"], "unit": "GB/s",
                             "peak": json.loads((ROOT / "MEASURED_PEAKS.json").read_text()).get("hbm_gbs", 6650.0)
                             if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0,
                             "traffic": None, "note": "algorithmic input + output bytes over the device-resident step time"},
                "parity_check": q["parity"]}
        line["roofline"]["frac"] = line["roofline"]["achieved"] / line["roofline"]["peak"]
        print(json.dumps(line), flush=True)
        return
    nw_total = args.nwavel if args.nwavel > 0 else c["nwavel"]
    with_wf = c["wf"]
    starts, counts = zip(*[wavelength_block(nw_total, r, world) for r in range(world)])
    start, count = starts[rank], counts[rank]
    sk._lib.check(sk._lib.lib().sk_b200_set_device(local_rank), "set_device")
    # every rank builds only its own block of the global spectrum (same formulae as the full scenario) and places it at
    # its offset of full-size caller arrays
    sc = build_scenario(args.config, nw_total, (start, count))
    _, geo, view, eng, atm = make_engine(sc, args.config, total_wavelengths=nw_total if world > 1 else None, start=start if world > 1 else 0)
    eng.set_workspace_gb(args.workspace_gb)
    nloc, nleg, nlos, nw = sc.nloc, sc.leg_coeff.shape[0], sc.nlos, sc.nwavel
    lib = sk._lib.lib()

    # caller-side result arrays: rank 0 is the caller that ends up holding the whole spectrum
    shapes = eng.result_shapes(atm)
    shared = None
    if world > 1:
        shared = SharedResult(shapes, rank, f"bench_{os.environ.get('MASTER_PORT', '0')}", barrier)
        buffers = shared.arrays
        out_pinned = shared.pinned
    else:
        buffers = {k: sk._lib.pinned_empty(s) for k, s in shapes.items()}
        out_pinned = True
    # page-lock the caller-side input buffers of this rank's block (the C ABI takes plain host pointers)
    in_arrays = [atm.storage.ssa, atm.storage.total_extinction, atm.storage.leg_coeff, atm.storage.solar_irradiance,
                 atm.surface.albedo]
    registered = []
    for arr in in_arrays:
        if lib.sk_b200_host_register(arr.ctypes.data, arr.nbytes) == 0:
            registered.append(arr)
    in_pinned = len(registered) == len(in_arrays)

    per_w_in = 8 * (nloc * (2 + nleg) + 2)
    nwf_out = 0
    if with_wf:
        for mp in sc.mappings.values():
            per_w_in += 8 * sum(int(np.prod(v.shape[:-1])) for v in mp.values() if isinstance(v, np.ndarray))
            nwf_out += sc.nloc
        nwf_out += 1
        per_w_in += 8  # d_brdf
    per_w_out = 8 * nlos * (1 + nwf_out)
    h2d_bytes, d2h_bytes = per_w_in * nw_total, per_w_out * nw_total

    # ---- device-resident timing (value): inputs staged once, K solves, CUDA events inside the library
    eng.stage(atm, start if world > 1 else 0, count, buffers=buffers)
    info = eng.info()
    m_list = list(range(info["num_azimuth"]))
    for _ in range(args.warmup):
        eng.solve_staged()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    dev_ms, per_kernel, launches = 0.0, {}, 0
    for _ in range(args.steps):
        eng.solve_staged()
        t = eng.timings_ms()
        dev_ms += t["kernels_total"]
        for k in ("optics", "layer", "bvp", "radiance", "wf_adjoint", "wf_layer", "wf_chain", "wf_map"):
            per_kernel[k] = per_kernel.get(k, 0.0) + t.get(k, 0.0)
        launches += eng.kernel_launches()
    barrier()
    wall_ms = reduce_max((time.perf_counter() - t0) * 1e3)
    clocks = sampler.stop()
    dev_ms = reduce_max(dev_ms)
    units_per_step = float(nw_total * nlos)
    ms_per_step = dev_ms / args.steps
    value = units_per_step / (ms_per_step * 1e-3)
    launches_all = int(reduce_sum(launches))

    # ---- end to end (e2e): host buffers in, all results in rank 0's host arrays, inside the timed region
    gather = None
    if world == 1:
        for _ in range(min(args.warmup, 2)):
            eng.calculate_radiance(atm, buffers=buffers)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            result = eng.calculate_radiance(atm, buffers=buffers)
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
        t_last = eng.timings_ms()
        breakdown = {"h2d_ms": t_last["h2d"], "kernels_ms": t_last["kernels_total"], "d2h_ms": t_last["d2h"]}
        breakdown["host_ms"] = e2e_ms - sum(breakdown.values())
    else:
        init_nccl_comm(rank, world)
        # (a) NCCL gather inside the library: H2D block -> solve -> ncclSend/Recv onto GPU 0 -> one D2H on rank 0
        def step_nccl():
            eng.stage(atm, start, count, buffers=buffers)
            eng.solve_staged()
            return eng.gather(starts, counts, nw_total, root=0)
        # (b) direct: every rank's block call writes straight into the caller's shared page-locked arrays
        def step_direct():
            return eng.calculate_radiance(atm, buffers=buffers, wavelength_block=(start, count))
        gather = {}
        for name, fn in (("nccl", step_nccl), ("direct", step_direct)):
            for _ in range(min(args.warmup, 2)):
                fn()
            barrier()
            t0 = time.perf_counter()
            nccl_ms = d2h_ms = 0.0
            for _ in range(args.steps):
                r = fn()
                if name == "nccl":
                    nccl_ms += r[1][0]
                    d2h_ms += r[1][1]
            barrier()
            ms = reduce_max((time.perf_counter() - t0) * 1e3) / args.steps
            gather[name] = {"ms_per_step": ms, "value": units_per_step / (ms * 1e-3)}
            if name == "nccl":
                gather[name]["nccl_exchange_ms_max"] = reduce_max(nccl_ms / args.steps)
                gather[name]["root_d2h_ms"] = reduce_max(d2h_ms / args.steps)
                gather[name]["bytes_over_nvlink"] = int(per_w_out * (nw_total - counts[0]))
        best = min(gather, key=lambda k: gather[k]["ms_per_step"])
        e2e_ms = gather[best]["ms_per_step"]
        gather["used_for_e2e"] = best
        gather["limiter"] = ("nccl variant: all results cross GPU 0's single PCIe link after the NVLink exchange; direct variant: "
                             "every rank's PCIe link carries only its own block")
        result = {k.replace("wf:", "").replace("surf:", ""): (v if not k.startswith("surf:") else v[0]) for k, v in buffers.items()}
        breakdown = None
    e2e_value = units_per_step / (e2e_ms * 1e-3)

    # ---- e2e from pageable caller buffers (what a numpy / ndarray caller hands over without sk_b200_host_alloc)
    e2e_pageable = None
    if world == 1:
        for arr in registered:
            lib.sk_b200_host_unregister(arr.ctypes.data)
        registered = []
        pageable = {k: np.empty(s) for k, s in shapes.items()}
        eng.calculate_radiance(atm, buffers=pageable)
        t0 = time.perf_counter()
        nrep = max(2, min(args.steps, 5))
        for _ in range(nrep):
            eng.calculate_radiance(atm, buffers=pageable)
        ms = (time.perf_counter() - t0) * 1e3 / nrep
        e2e_pageable = {"value": units_per_step / (ms * 1e-3), "ms_per_step": ms,
                        "note": "inputs and outputs in pageable host memory (no cudaHostRegister / sk_b200_host_alloc)"}
        del pageable

    # ---- parity of the timed outputs against the oracle (rank 0's block)
    threads = os.cpu_count() or 1
    parity = None
    if rank == 0:
        res_local = result
        if world > 1:   # rank 0's block of the shared full-spectrum arrays
            res_local = {k: (v[start:start + count] if v.shape[0] == nw_total else v[:, start:start + count]) for k, v in result.items()}
        parity = parity_sample(sc, args.config, res_local, 12, threads)
    ok = reduce_max(0.0 if (parity is None or parity["ok"]) else 1.0)
    if ok != 0.0:
        raise SystemExit(f"bench.py: timed outputs disagree with the oracle: {parity}")

    # ---- roofline of the dominant kernel (FP64 pipe; peak measured live by a DFMA micro-benchmark)
    reuse = bool(lib.sk_b200_adjoint_reuses_factors(sc.nstr // 2, nlos))
    fm = flop_model(sc.nstr, sc.nloc - 1, nlos, m_list, ngroups=1 if with_wf else 0, adjoint_refactor=not reuse)
    fast = sc.nstr in (4, 8, 16) and os.environ.get("SK_B200_GENERIC", "0") != "1"
    kernel_names = {
        "layer": "k_eig_setup + k_eig_jacobi + k_layer_post (+ k_los_atten)" if fast else "k_layer_solve",
        "bvp": "k_bvp_v2 (blocked elimination)" if sc.nstr <= 16 else "k_bvp",
        "wf_adjoint": ("k_bvp_tsolve" if reuse else "k_bvp_adjoint_v2") if sc.nstr <= 16 else "k_bvp_adjoint",
        "wf_layer": "k_wf_layer_fast" if fast else "k_wf_layer",
    }
    launches_per_chunk = {"layer": 4 if fast else 1, "bvp": 1, "wf_adjoint": 1, "wf_layer": 1}
    chunk = info["chunk_wavelengths"]
    nchunks = int(np.ceil(nw / chunk))
    try:
        fp64_peak = float(lib.sk_b200_measure_fp64_tflops())
    except Exception:
        fp64_peak = None
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    step_kernel_ms = max(sum(per_kernel.values()), 1e-12)

    def kernel_roofline(k):
        ms = per_kernel.get(k, 0.0)
        if ms <= 0.0:
            return None
        n_launch = args.steps * nchunks * launches_per_chunk[k]
        tf = fm[k] * nw * args.steps / (ms * 1e-3) / 1e12   # algorithmic flops of the group / its device time
        return {"kernel": kernel_names[k], "achieved": tf, "frac": (tf / fp64_peak) if fp64_peak else None,
                "share_of_step": ms / step_kernel_ms, "avg_launch_ms": ms / max(n_launch, 1),
                "flops_per_launch": fm[k] * (nw / nchunks) / launches_per_chunk[k]}

    per_k = {k: kernel_roofline(k) for k in ("layer", "bvp", "wf_adjoint", "wf_layer")}
    per_k = {k: v for k, v in per_k.items() if v}
    if reuse and "wf_adjoint" in per_k:
        by = tsolve_bytes_model(sc.nstr, sc.nloc - 1, nlos, len(m_list))
        gbs = by * nw * args.steps / (per_kernel["wf_adjoint"] * 1e-3) / 1e9
        per_k["wf_adjoint"]["hbm"] = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                                      "frac": gbs / hbm_peak, "bytes_per_wavelength": by}
    dom = max(per_k, key=lambda k: per_k[k]["share_of_step"])
    total_flops = fm["total_wf"] if with_wf else fm["total"]
    alg_bytes = bytes_model(nloc, nleg, nlos, nwf_out, 1 if with_wf else 0)
    roofline = {
        "bound": "fp64", "kernel": per_k[dom]["kernel"], "achieved": per_k[dom]["achieved"], "peak": fp64_peak,
        "unit": "TFLOP/s", "frac": per_k[dom]["frac"],
        "traffic": (NCU_DRAM_BYTES_PER_WAVELENGTH[dom] * (nw / nchunks) / launches_per_chunk[dom]
                    if (sc.nstr, sc.nloc - 1, nlos, with_wf) == (16, 100, 10, True) and dom in NCU_DRAM_BYTES_PER_WAVELENGTH else None),
        "traffic_source": NCU_SOURCE + " (dram bytes per wavelength of the profiled launch x wavelengths per launch)",
        "peak_source": "DFMA micro-benchmark run inside this bench (MEASURED_PEAKS.json has no FP64 figure)",
        "share_of_step": per_k[dom]["share_of_step"], "avg_launch_ms": per_k[dom]["avg_launch_ms"],
        "flops_per_launch": per_k[dom]["flops_per_launch"],
        "kernels": per_k,
        "whole_step": {"flops_per_wavelength": total_flops,
                       "achieved_tflops": total_flops * nw / (ms_per_step * 1e-3) / 1e12,
                       "frac": (total_flops * nw / (ms_per_step * 1e-3) / 1e12 / fp64_peak) if fp64_peak else None},
        "hbm_view": {"bound": "hbm", "achieved": alg_bytes * nw / (ms_per_step * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                     "algorithmic_bytes_per_wavelength": alg_bytes,
                     "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"},
    }
    roofline["hbm_view"]["frac"] = roofline["hbm_view"]["achieved"] / hbm_peak

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"{c['label']}: {c['text']}", "total_wavelengths": nw_total,
                   "wavelengths_per_gpu": [int(v) for v in counts], "azimuth_orders": len(m_list), "chunk_wavelengths": chunk,
                   "sharding": "contiguous wavelength blocks, one process per GPU, no exchange inside the solve",
                   "l2": "inputs+workspace per step far exceed the 126 MB L2 (no flush needed)",
                   "wall_ms_per_step": wall_ms / args.steps},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                "ms_per_step": e2e_ms, "pinned": bool(in_pinned and out_pinned), "breakdown": breakdown, "gather": gather,
                "pageable": e2e_pageable},
        "gpu_launches": launches_all,
        "kernel_ms_per_step": {k: v / args.steps for k, v in per_kernel.items()},
        "roofline": roofline,
        "parity_check": parity,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        scd = build_scenario_decimated(args.config, nw_total, args.cpu_sample) if nw_total > 4 * args.cpu_sample else sc
        cpu_arm(scd, args.config, min(args.cpu_sample, 2 * threads), threads)
        v, dt, desc = cpu_arm(scd, args.config, args.cpu_sample, threads)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc}
    if rank == 0 and world == 1 and not args.no_other_configs:
        del result, buffers
        others = {}
        for name in ("c2", "c1", "c3", "c4"):
            if name == args.config:
                continue
            try:
                others[name] = quick_config(name, 2, 1, local_rank, threads)
            except Exception as ex:  # a workload whose path is not built yet must not hide the headline
                others[name] = {"error": f"{type(ex).__name__}: {ex}"}
        line["other_configs"] = others
    if rank == 0:
        print(json.dumps(line), flush=True)
    for arr in registered:
        lib.sk_b200_host_unregister(arr.ctypes.data)
    if shared is not None:
        lib.sk_b200_comm_destroy()
        shared.close(barrier)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
